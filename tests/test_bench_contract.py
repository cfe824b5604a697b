"""bench.py contract pieces that can run without a GPU: the reference (CPU) arm prints ONE JSON line with the keys the
driver reads, and the agent-chunking arithmetic cuts a batch into equal whole-tile parts."""
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--agents", "32"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "agent-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("agent-steps/sec") and d["value"] > 0 and d["n_gpus"] == 1
    for k in ("steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["config"]["workload"].startswith("configs[2]")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_equal_part_chunking():
    def parts(B, cap):
        n = max(1, -(-B // cap))
        return n, min(B, -(-(-(-B // n)) // 128) * 128)
    assert parts(1_000_000, 378_880) == (3, 333_440)
    assert parts(10_000, 378_880) == (1, 10_000)
    assert parts(378_880, 378_880) == (1, 378_880)
    n, c = parts(8_000_000, 378_880)
    assert n == 22 and c % 128 == 0 and c <= 378_880 and n * c >= 8_000_000


def test_chunk_bounds_cover_the_batch_in_whole_waves():
    """bench.chunk_bounds: parts cover [0, B) without gaps, none larger than the cap, all but the last a whole number of waves"""
    import bench
    for B, cap in [(1_000_000, 265_216), (125_000, 265_216), (700_000, 265_216), (10_000, 265_216), (8_000_000, 265_216), (300_001, 100_000)]:
        parts = bench.chunk_bounds(B, cap, 296)
        assert parts[0][0] == 0 and parts[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(0 < e - s <= cap + 296 * 128 for s, e in parts)
        if len(parts) > 1 and -(-(-(-B // 128)) // 296) >= len(parts):
            assert all((e - s) % (296 * 128) == 0 for s, e in parts[:-1])


def test_continuous_rk4_mode_is_configs4_as_one_solve():
    """--workload c5 --adjoint-mode continuous-rk4: t = [0, 24] with step_size 0.25 = 96 grid steps, rk4, all agents in one solve;
    an agent-step of that line is one grid step (not one output interval)."""
    import types
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import bench
    args = types.SimpleNamespace(workload="c5", agents=0, solver="", adjoint_mode="continuous-rk4")
    cfg = bench._config_for(args)
    assert cfg["method"] == "rk4" and cfg["T"] == 2 and cfg["step_size"] == 0.25 and cfg["grid_steps"] == 96 and cfg["B"] == 8_000_000
    assert cfg["adjoint"] and "configs[4]" in cfg["name"]
    _, _, _, t = bench.make_inputs(dict(cfg, B=4))
    assert t.tolist() == [0.0, 24.0]
    from ananke_abm_b200.adjoint_tc import step_grid
    assert len(step_grid(0.0, 24.0, cfg["step_size"])) - 1 == cfg["grid_steps"]
    args = types.SimpleNamespace(workload="c5", agents=1000, solver="", adjoint_mode="discrete")
    cfg = bench._config_for(args)
    assert cfg["method"] == "dopri5" and cfg["T"] == 97 and cfg["B"] == 1000 and "grid_steps" not in cfg


def test_fused_step_coefficients_reproduce_the_stage_recursion():
    """adjoint_tc.fused_step_coefficients: the linear form of an augmented 3/8-rule step over (a0, gx'_0..gx'_3) against the stage
    recursion written out directly on scalars (a_p, a_v) with an arbitrary linear 'Jacobian' per stage"""
    import random
    from ananke_abm_b200 import adjoint_tc
    from ananke_abm_b200.stage import RK38
    rnd = random.Random(3)
    h = -0.37
    a0p, a0v = rnd.uniform(-1, 1), rnd.uniform(-1, 1)
    Jp = [rnd.uniform(-1, 1) for _ in range(4)]      # gx_s.p = Jp[s] * a_v,s ; gx_s.v = Jv[s] * a_v,s
    Jv = [rnd.uniform(-1, 1) for _ in range(4)]
    # direct recursion: ka_s = -[gx_s.p, a_p,s + gx_s.v]
    ka, gxs = [], []
    for s in range(4):
        ap = a0p + h * sum(b * ka[j][0] for j, b in enumerate(RK38.beta[s]))
        av = a0v + h * sum(b * ka[j][1] for j, b in enumerate(RK38.beta[s]))
        gx = (Jp[s] * av, Jv[s] * av)
        gxs.append((gx, av))
        ka.append((-gx[0], -(ap + gx[1])))
    a1p = a0p + h * sum(RK38.b[s] * ka[s][0] for s in range(4))
    a1v = a0v + h * sum(RK38.b[s] * ka[s][1] for s in range(4))
    # linear form
    c, cpa, cva, dp, dv, w = adjoint_tc.fused_step_coefficients(h)
    gxp = []
    for s in range(4):
        up = cva[s] * a0v + cpa[s] * a0p + sum(dp[s][i] * gxp[i][0] + dv[s][i] * gxp[i][1] for i in range(s))
        assert abs(up - c[s] * gxs[s][1]) < 1e-12                       # upstream_s == c_s a_v,s
        gxp.append((Jp[s] * up, Jv[s] * up))                            # the scaled product c_s gx_s
    b1p = a0p + sum(g[0] for g in gxp)
    b1v = a0v - h * a0p + sum(w[i] * gxp[i][0] + gxp[i][1] for i in range(4))
    assert abs(b1p - a1p) < 1e-12 and abs(b1v - a1v) < 1e-12
