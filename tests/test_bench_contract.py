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
