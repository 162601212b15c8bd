"""bench.py's reference arm runs on a CPU box: check that it prints exactly one JSON line with the contract's keys.
(The GPU arm prints the same keys plus roofline / clocks; it needs a B200.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


@pytest.mark.parametrize("workload", ["c1", "target", "c2", "c5"])
def test_reference_arm_prints_one_contract_line(workload):
    rows = {"c1": "20000", "target": "40960", "c2": "4096", "c5": "30720"}[workload]
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--rows", rows,
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), sorted(REQUIRED - set(d))
    assert d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["config"]["workload"].startswith(workload)
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # both arms print the same `config` object (bench_workloads.Workload.config + the L2 note)
    for key in ("workload", "rows", "dim", "k", "nq", "chunk_size", "metric", "filter", "vec_filter", "planted_rows", "l2"):
        assert key in d["config"], key
    assert "thread count pinned explicitly" in d["cpu_baseline"]["sample"]


def test_reference_arm_ignores_omp_num_threads():
    """torchrun exports OMP_NUM_THREADS=1; the MetaStore reference arm must still use every usable core."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "target", "--rows", "40960",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_reference_arm_does_not_import_the_product():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'c1', '--rows', '5000', '--steps', '1', '--warmup', '0'];"
            "runpy.run_path('bench.py', run_name='__main__'); assert 'otters_b200' not in sys.modules, 'reference arm imported otters_b200'")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]


def test_workload_specs_are_consistent():
    """bench_workloads: the NumPy restatement of each filter agrees with the oracle's row mask on the same columns."""
    sys.path.insert(0, ROOT)
    import numpy as np

    import bench_workloads as bw
    from oracle import oracle as ora

    for name in ("target", "c5"):
        wl = bw.Workload(name, 50_000)
        cols = wl.columns(np.arange(wl.rows))
        idx = {c.name(): i for i, c in enumerate(cols)}
        fp = ora.FilterPack([[(idx[n], op, kind, val) for n, op, kind, val in cl] for cl in wl.clauses()])
        ost = ora.MetaStore(np.ones((wl.rows, 1), np.float32), cols, wl.chunk)
        assert np.array_equal(ost.row_mask(fp).astype(bool), wl.row_mask(cols)), name
        assert 0.2 < wl.row_mask(cols).mean() < 0.7
    # block-cyclic maps: global -> local inverts local -> global
    g = bw.cyclic_global_rows(10_000, 128, 4, 3)
    mine, loc = bw.global_to_local(np.arange(10_000), 128, 4, 3)
    assert np.array_equal(np.nonzero(mine)[0], g) and np.array_equal(loc, np.arange(len(g)))


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_scores_the_rounded_rows_for_bf16_stores():
    """--vector-format bf16: the config says so in both arms and the CPU arm works on f32(bf16(x)) rows (Workload.stored)."""
    sys.path.insert(0, ROOT)
    import numpy as np

    import bench_workloads as bw
    from oracle import oracle as ora

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "target", "--rows", "40960",
                          "--vector-format", "bf16", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["config"]["vector_format"] == "bf16" and d["value"] > 0
    wl = bw.Workload("c1", 1000, "bf16")
    v = bw.synth_fill_np(0, 64, wl.dim, bw.DATA_SEED)
    assert np.array_equal(wl.stored(v).view(np.uint32), ora.round_bf16(v).view(np.uint32))
    assert bw.Workload("c1", 1000).stored(v) is v and bw.Workload("c1", 1000).config()["vector_format"] == "f32"
