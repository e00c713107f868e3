"""Run bench.py's control flow WITHOUT a GPU: CUDA streams/events, the device models and the search Session are replaced by
host stubs, NCCL by gloo.  Only the sequencing is exercised (phases, collectives, JSON line, exit path) -- never a bench
value.  Used by tests/test_bench_plumbing.py in a subprocess (the patches must not leak into other tests)."""
import contextlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

_real_device = torch.device
torch.device = lambda *a, **k: _real_device("cpu") if a and a[0] == "cuda" else _real_device(*a, **k)


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class _Stream:
    def __init__(self, device=None):
        pass

    def wait_event(self, e):
        pass

    def synchronize(self):
        pass


torch.cuda.set_device = lambda d: None
torch.cuda.Stream = _Stream
torch.cuda.Event = _Event
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.synchronize = lambda d=None: None
torch.cuda.current_stream = lambda d=None: _Stream()
torch.cuda.empty_cache = lambda: None

import atspeed_b200.engine as E  # noqa: E402
import bench  # noqa: E402


class _Model:
    def __init__(self, spec, W, dev):
        self.spec, self.device = spec, dev


class _Trie:
    def __init__(self, csr, dev):
        pass


class _Session:
    def __init__(self, tdm, ddm, trie, K, *a, **k):
        self.K = K

    def _stats(self):
        return {"n_run": 3, "total_accept_steps": 1, "target_forwards": 4, "draft_forwards": 6, "kernel_launches": 1000}

    def bssd_batch_device(self, cat, lens, gamma, tok, sc):
        return [self._stats() for _ in lens]

    def bssd_device(self, p, gamma, tok, sc):
        return self._stats()

    def _host(self):
        return dict(self._stats(), tokens=np.zeros((self.K, 4), np.int32), scores=np.zeros(self.K, np.float32))

    def bssd_batch(self, prompts, gamma):
        return [self._host() for _ in prompts]

    def bssd(self, p, gamma):
        return self._host()

    def set_shared_prefix(self, ids):
        return len(ids)

    def profile(self, on):
        pass

    def profile_read(self):
        names = ("gemm", "attention", "rowwise", "topk", "beam", "kvgather")
        return {n: {"ms": 1.0 + i, "launches": 10, "bytes": 1e9 if i in (0, 3) else 0.0, "flops": 1e12 if i == 0 else 0.0}
                for i, n in enumerate(names)}


class _Sampler(bench.ClockSampler):
    def __enter__(self):
        return self


class _Builder:
    def __init__(self, ds, dev):
        self.ds = ds

    def build(self, users):
        ps = [self.ds.prompt_ids(u) for u in users]
        return torch.tensor([t for p in ps for t in p], dtype=torch.int32), [len(p) for p in ps]


import atspeed_b200.prompts as P  # noqa: E402
P.DevicePromptBuilder = _Builder
E.DeviceModel, E.DeviceTrie, E.Session = _Model, _Trie, _Session
bench.gpu_weights = lambda spec, seed, dev: {}
bench.ClockSampler = _Sampler

if int(os.environ.get("WORLD_SIZE", 1)) > 1:
    import torch.distributed as dist
    _real_init = dist.init_process_group
    dist.init_process_group = lambda backend, device_id=None, timeout=None: _real_init("gloo", timeout=timeout)

if __name__ == "__main__":
    import atexit
    atexit.register(lambda: sys.stderr.write("[stub] atexit hook ran\n"))
    sys.argv = ["bench.py"] + sys.argv[1:]
    sys.exit(bench.main())
