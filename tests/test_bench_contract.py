"""CPU: the parts of bench.py's contract that need no GPU -- both arms describe the workload with the same `config`
dictionary, the reference arm runs on the host alone and prints one JSON line with the keys the driver reads."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_arm_config_is_computable_without_a_mesh_and_matches_the_known_sizes():
    import bench

    a = argparse.Namespace(scaling="weak", nex="128x128x64", halo="p2p")
    c1, c8 = bench.arm_config(a, 1), bench.arm_config(a, 8)
    assert c1["elements_total"] == 1048576 and c1["nglob_per_gpu"] == 513 * 513 * 257 and c1["halo"] == "none"
    assert c8["elements_total"] == 8 * 1048576 and c8["elements_per_gpu"] == 1048576 and c8["partitions"].startswith("4x2")
    assert c1["dt"] == c8["dt"]  # weak scaling keeps the element size
    a.scaling = "strong"
    s1, s8 = bench.arm_config(a, 1), bench.arm_config(a, 8)
    assert s1["elements_total"] == s8["elements_total"] == 256 * 256 * 128 and s8["elements_per_gpu"] == 1048576
    assert "configs[4]" in s8["workload"] and abs(s8["dt"] - 0.5 * c1["dt"]) < 1e-12


def test_reference_arm_prints_the_contract_line_with_the_b200_arms_config():
    import bench

    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                          "--nex", "16x16x8"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=600,
                         env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert res.returncode == 0
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1  # ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    a = argparse.Namespace(scaling="weak", nex="16x16x8", halo="p2p")
    assert d["config"] == bench.arm_config(a, 2)
