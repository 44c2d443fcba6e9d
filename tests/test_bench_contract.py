"""CPU suite, part 5: bench.py's reference arm runs without a GPU and prints one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-pairs", "2", "--modality", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "gan_train_step_pairs_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "workload" in line["config"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_algorithmic_work_matches_the_survey_table():
    sys.path.insert(0, ROOT)
    import bench
    for D, gf, mb in ((400, 1.456, 46.7), (1200, 2.336, 80.8), (3632, 5.012, 184.4), (12032, 14.25, 542.5)):
        flops, nbytes, N_D, N_G = bench.algo_work(D, 50)
        assert abs(flops / 1e9 - gf) < 0.01 * gf and abs(nbytes / 1e6 - mb) < 0.01 * mb
        assert N_D == 1000 * D + 753756 and N_G == 501 * D + 302000
