"""CPU tests of bench.py's process plumbing (no GPU work): the watchdog that turns a stalled run into stack dumps plus
whatever was measured, the conservative multi-GPU environment defaults, and the reference arm's rank gating."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, env=None, timeout=120):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=e, capture_output=True, text=True, timeout=timeout)


def test_watchdog_prints_partial_result_and_stacks():
    r = _run("import time, bench\n"
             "bench.STATE['partial'] = {'metric': 'topk_recs_per_sec', 'value': 1.5}\n"
             "bench.log('host-buffer pass: timed region')\n"
             "bench.arm_watchdog(100, 1)\n"
             "time.sleep(30)\n")
    assert r.returncode == 17, r.stderr          # a stalled run is a failed run, even with a partial line
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["value"] == 1.5 and "host-buffer pass" in line["incomplete"]
    assert "watchdog [r0]: no progress for 1 s" in r.stderr and "File" in r.stderr      # faulthandler stack dump


def test_watchdog_without_a_measurement_fails_the_run():
    r = _run("import time, bench\nbench.arm_watchdog(1, 100)\ntime.sleep(30)\n")
    assert r.returncode == 17 and r.stdout.strip() == "" and "no result after 1 s" in r.stderr


def test_watchdog_other_ranks_stay_silent():
    r = _run("import time, bench\nbench.STATE['partial'] = {'value': 2}\nbench.arm_watchdog(1, 100)\ntime.sleep(30)\n",
             env={"RANK": "3"})
    assert r.returncode == 17 and r.stdout.strip() == "" and "watchdog [r3]" in r.stderr


def test_watchdog_after_the_result_line_is_a_success():
    r = _run("import time, bench\nbench.STATE['printed'] = True\nbench.arm_watchdog(1, 100)\ntime.sleep(30)\n")
    assert r.returncode == 0 and "no result after 1 s" in r.stderr


def test_multi_gpu_env_defaults_do_not_override_the_caller_and_never_touch_the_kernels():
    r = _run("import os, bench\n"
             "os.environ.pop('ATSPEED_GEMM_2CTA', None); os.environ.pop('NCCL_NVLS_ENABLE', None)\n"
             "bench.multi_gpu_env(1)\n"
             "assert 'ATSPEED_GEMM_2CTA' not in os.environ and 'NCCL_NVLS_ENABLE' not in os.environ\n"
             "bench.multi_gpu_env(8)\n"
             "assert 'ATSPEED_GEMM_2CTA' not in os.environ, 'every N runs the same GEMM configuration'\n"
             "assert os.environ['NCCL_NVLS_ENABLE'] == '0'\n"
             "os.environ['NCCL_NVLS_ENABLE'] = '1'\n"
             "bench.multi_gpu_env(8)\n"
             "assert os.environ['NCCL_NVLS_ENABLE'] == '1'\n")
    assert r.returncode == 0, r.stderr


def test_reference_arm_runs_on_rank_0_only():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), capture_output=True,
                       text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


STUB = os.path.join(ROOT, "tests", "helpers", "bench_flow_stub.py")
KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "kernel_groups", "cpu_baseline")


def test_bench_control_flow_single_process():
    """bench.py's own sequencing (phases, JSON keys) with the CUDA session stubbed out: catches a broken bench line on CPU."""
    for extra in ([], ["--cohort", "1", "--lanes", "2"]):
        r = subprocess.run([sys.executable, STUB, "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--hf-baseline-users", "0",
                            "--check-users", "0"]
                           + extra, cwd=ROOT, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        line = json.loads(r.stdout.strip().splitlines()[-1])
        assert all(k in line for k in KEYS), [k for k in KEYS if k not in line]
        assert line["n_gpus"] == 1 and line["config"]["gemm_pair_kernel"] is True and "incomplete" not in line
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
        assert "device pass: timed region" in r.stderr and "done" in r.stderr and "[stub] atexit hook ran" in r.stderr


def test_bench_control_flow_two_ranks_gloo():
    """The N = 2 launch exactly as the driver does it (torchrun), NCCL replaced by gloo: every rank walks the same collectives,
    rank 0 alone prints ONE line, the kernel configuration is the one of N = 1, and every process exits 0 through the
    interpreter's normal exit (atexit hooks run: the driver's native-library hook depends on it)."""
    env = {k: v for k, v in os.environ.items() if k not in ("ATSPEED_GEMM_2CTA", "NCCL_NVLS_ENABLE")}
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29677", STUB, "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["config"]["gemm_pair_kernel"] is True and line["cpu_baseline"] is None
    assert line["gpu_launches"] > 0 and line["scaling"] == "weak"
    for rank in (0, 1):
        assert f"[bench r{rank} " in r.stderr and "all-gather done" in r.stderr
    assert r.stderr.count("[stub] atexit hook ran") == 2
