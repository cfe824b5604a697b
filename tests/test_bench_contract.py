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
