#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the atspeed_b200 hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): LLaMA-7B-shape target + LLaMA-68M-shape draft (random-init bf16
weights), AtSpeed-S strict top-K verify, Beauty test users, K=10 target beams, N=40 draft beams, gamma=3,
4 new tokens, strict item trie.  A "step" = one batch of `--users-per-step` users, each taken through the
whole speculative beam search (reference code/beamSD.py:458-542).  metric = ranked top-K lists produced
per second over the whole job ("users/s").  By default the users go through cohort sessions: 16 searches in flight per lane
share every forward (<= 512 tokens), 2 lanes per GPU, and the K/V of the opening tokens all prompts have in common (the
instruction template, 39 tokens) is computed once per session outside the timed region (`config.shared_prompt_prefix_tokens`;
`--no-shared-prefix` switches that off, `--cohort 1 --lanes 1` is the reference's one search at a time).

  value : history item ids resident in HBM, prompts built on the device, results left in HBM (atspeed_bssd_batch_device);
          timed with CUDA events on the launching stream between barriers, max over ranks.
  e2e   : same users through the host-buffer C ABI call (atspeed_bssd_batch): prompt ids copied host->device and
          ranked lists + scores device->host inside the timed region.
  roofline : the dominant kernel (the tcgen05 GEMM): algorithmic bytes / CUDA-event duration per launch,
          measured in a second pass over the same users with per-launch events enabled (the first pass stays
          unperturbed), against MEASURED_PEAKS.json.
  cpu_baseline : oracle port (oracle/bssd_ref.py + oracle/llama_ref.py) on the host cores, bounded sample.

Ranks shard users (rank r takes users r, r+W, ...); the only collective is one all-gather of the ranked
lists per step.  `--impl reference` times the CPU oracle port alone (rank 0 only).

Also on the line: `pass_consistency` (the host-buffer pass re-runs the device-resident pass's last step: the ranked lists
must be identical), `kernel_groups` (share of a user's GPU time and algorithmic GB/s per kernel family), `hf_gpu_baseline`
(transformers generate(num_beams=K) on the same GPU, N = 1: the north star's ">= 2x HF beam search" comparison, reference
code/inference.py:177-181 `speedupTF`).  stderr carries phase breadcrumbs; a watchdog ends a stalled run with every thread's
stack, the partial result (`"incomplete"`) and a NON-ZERO exit code, see DESIGN.md section 6.  Every N runs the same kernel
configuration (the CTA-pair GEMM serves forwards of > 256 tokens at any N unless ATSPEED_GEMM_2CTA=0; `config.gemm_pair_kernel`
says which).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_START = time.perf_counter()
STATE = {"rank": int(os.environ.get("RANK", 0)), "phase": "start", "t_phase": T_START, "partial": None}


def log(phase):
    """Progress breadcrumb on stderr (rank, seconds since start, phase): a stalled multi-GPU run shows where it stopped."""
    STATE["phase"], STATE["t_phase"] = phase, time.perf_counter()
    sys.stderr.write("[bench r%d +%.1fs] %s\n" % (STATE["rank"], time.perf_counter() - T_START, phase))
    sys.stderr.flush()

SHAPES = {"7b": dict(hidden=4096, n_layers=32, n_heads=32, mlp=11008),
          "68m": dict(hidden=768, n_layers=2, n_heads=12, mlp=3072),
          "small": dict(hidden=256, n_layers=2, n_heads=4, mlp=512),
          "small_draft": dict(hidden=128, n_layers=1, n_heads=2, mlp=256)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="atspeed", choices=["atspeed", "reference"])
    ap.add_argument("--users-per-step", type=int, default=96)
    ap.add_argument("--cohort", type=int, default=16,
                    help="users whose trees share each forward (atspeed_bssd_batch; <= 512 tokens per forward); 1 = one "
                         "search per forward as the reference")
    ap.add_argument("--dataset", default="beauty")
    ap.add_argument("--K", type=int, default=10)
    ap.add_argument("--N", type=int, default=40)
    ap.add_argument("--gamma", type=int, default=3)
    ap.add_argument("--target", default="7b")
    ap.add_argument("--draft", default="68m",
                    help="draft shape; 'corr<n>' (e.g. corr24) = a CORRELATED draft: the target's first n layers, embedding and "
                         "lm_head plus 3 %% noise.  Independent random weights accept only the structurally forced last step; "
                         "this exercises real accepted steps at the target's shape (a random-init net needs most of its layers "
                         "to agree with itself: corr2 accepts nothing extra, measured)")
    ap.add_argument("--check-users", type=int, default=3,
                    help="parity record at the benchmark shape: this many users also go through oracle/bssd_ref.py on the host "
                         "with the GPU's own weights (bf16 contract); ranked lists + accepted lengths compared (~10 s + 3.5 s/user on "
                         "32 cores; N = 1 only; 0 = skip)")
    ap.add_argument("--constraint", default="strict", choices=["strict", "positional"])
    ap.add_argument("--profile-users", type=int, default=4)
    ap.add_argument("--lanes", type=int, default=2,
                    help="independent searches in flight per GPU (each its own session + CUDA stream + host thread): one "
                         "user's latency-bound draft / verify phases overlap another's weight-streaming target forward")
    ap.add_argument("--cohort-tokens", type=int, default=512, help="most tokens one cohort forward packs (256..512)")
    ap.add_argument("--do-sample", action="store_true", help="AtSpeed-R relaxed acceptance (configs[2]) instead of AtSpeed-S")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shared-prefix", action="store_true",
                    help="cohort mode: do not share the K/V of the prompts' common opening tokens (the instruction template) between users")
    ap.add_argument("--emulate-shard", default=None, metavar="R/W",
                    help="diagnostic: on ONE GPU, process the user slice rank R of W would get (no collective), e.g. 0/8")
    ap.add_argument("--hf-baseline-users", type=int, default=12,
                    help="also time HF generate(num_beams=K) on the same GPU after 3 warm-up calls (N=1 only; 0 = skip)")
    return ap.parse_args()


def corr_layers(name):
    """'corr<n>' -> n (a correlated draft made of the target's first n layers), anything else -> 0."""
    return int(name[4:]) if name.startswith("corr") and name[4:].isdigit() else 0


def workload_name(a):
    mode = "AtSpeed-R relaxed acceptance (do_sample, top_k=50, T=1)" if a.do_sample else "AtSpeed-S strict top-K verify"
    draft = (f"correlated {corr_layers(a.draft)}-layer draft cut from the target (+3% noise)" if corr_layers(a.draft)
             else f"LLaMA-{a.draft}-shape draft")
    return (f"LLaMA-{a.target}-shape target + {draft}, {mode}, "
            f"{a.dataset} test users, {a.constraint} constraint, K={a.K} N={a.N} gamma={a.gamma} max_new_tokens=4, "
            f"{a.users_per_step} users/step/GPU, " +
            (f"cohorts of up to {a.cohort} users per forward (<={a.cohort_tokens} tokens), {a.lanes} cohorts in flight per GPU"
             if a.cohort > 1 else f"batch 1 per search (as the reference), {a.lanes} searches in flight per GPU"))


# ---------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------
def gpu_weights(spec, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    r = lambda *s: (torch.randn(*s, generator=g, device=device, dtype=torch.float32) * 0.02).to(torch.bfloat16)
    hd = spec.n_heads * spec.head_dim
    W = {"embed": r(spec.vocab, spec.hidden), "norm": torch.ones(spec.hidden, device=device, dtype=torch.bfloat16),
         "lm_head": r(spec.vocab, spec.hidden), "layers": []}
    for _ in range(spec.n_layers):
        W["layers"].append({"wq": r(hd, spec.hidden), "wk": r(hd, spec.hidden), "wv": r(hd, spec.hidden),
                            "wo": r(spec.hidden, hd), "wg": r(spec.mlp, spec.hidden), "wu": r(spec.mlp, spec.hidden),
                            "wd": r(spec.hidden, spec.mlp),
                            "ln1": torch.ones(spec.hidden, device=device, dtype=torch.bfloat16),
                            "ln2": torch.ones(spec.hidden, device=device, dtype=torch.bfloat16)})
    return W


def correlated_draft_weights(tdm, n_layers, noise, device):
    """The target's embedding, first `n_layers` layers, final norm and lm_head, each matrix perturbed by `noise` x its own
    standard deviation: a draft whose beams agree with the target often enough to exercise accepted steps (SURVEY hard part 5)."""
    g = torch.Generator(device=device).manual_seed(3)

    def pert(t):
        if t.dim() < 2:
            return t.clone()
        n = torch.randn(t.shape, generator=g, device=device, dtype=torch.float32)
        return (t.float() + noise * t.float().std() * n).to(torch.bfloat16)

    return {"embed": pert(tdm.embed), "norm": tdm.norm.clone(), "lm_head": pert(tdm.lm_head),
            "layers": [{k: pert(v) for k, v in ly.items()} for ly in tdm.layers[:n_layers]]}


def check_users_against_oracle(a, ds, fn, tdm, ddm, sess, users, prompts_host):
    """Parity record at the benchmark shape: the same users through oracle/bssd_ref.py on the host cores
    with the GPU's own weights (bf16 contract of oracle/llama_ref.py), single search each.  Strict mode only."""
    from oracle import bssd_ref
    from oracle import llama_ref as LR
    host_threads()
    models = []
    for dm in (tdm, ddm):
        sp = dm.spec
        sh = LR.LlamaShape(sp.vocab, sp.hidden, sp.n_layers, sp.n_heads, sp.mlp, sp.head_dim, sp.rope_theta, sp.eps)
        W = {"embed": dm.embed, "norm": dm.norm, "lm_head": dm.lm_head, "layers": dm.layers}
        models.append(LR.RefLlama(sh, W, "bf16"))
    tol = 6e-2
    rec = {"users": 0, "identical_ranked_lists": 0, "explained_by_near_tie": 0, "identical_accept_steps": 0, "unexplained": [],
           "tolerance": tol, "cpu_s_per_user": 0.0}
    t0 = time.perf_counter()
    for u in users:
        log("check-users: user %d through the oracle" % u)
        bssd_ref.GAP_LOG = []
        try:
            ref = bssd_ref.bssd(models[0], models[1], prompts_host[u], a.K, a.N, a.gamma, 4, fn)
            margin = min(bssd_ref.GAP_LOG) if bssd_ref.GAP_LOG else float("inf")
        finally:
            bssd_ref.GAP_LOG = None
        o = sess.bssd_batch([prompts_host[u]], a.gamma)[0] if a.cohort > 1 else sess.bssd(prompts_host[u], a.gamma)
        P = len(prompts_host[u])
        mine, theirs = o["tokens"][:, :4].tolist(), ref.sequences[:, P:].tolist()
        rec["users"] += 1
        rec["identical_accept_steps"] += int(list(o["accept_steps"]) == list(ref.accept_steps))
        if mine == theirs:
            rec["identical_ranked_lists"] += 1
        elif margin < tol:
            rec["explained_by_near_tie"] += 1
        else:
            rec["unexplained"].append({"user": int(u), "smallest_oracle_margin": margin})
    rec["cpu_s_per_user"] = (time.perf_counter() - t0) / max(1, rec["users"])
    return rec


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        if self.index is None:
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel, mode):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of THIS configuration
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from dram__bytes_read.sum + dram__bytes_write.sum per launch;
    entries are keyed "<kernel>/<mode>", mode = "cohort" | "single"); None when that kernel/mode was not captured -- a
    capture of another configuration is never substituted."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    j = json.load(open(p)).get("%s/%s" % (kernel, mode))
    if not j:
        return None, None
    return j["dram_bytes_per_launch"], j


def peaks():
    """(HBM GB/s, sustained bf16 TFLOP/s, source): the GEMM is timed inside a long step, so the sustained figure applies."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j["hbm_gbs"], j.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json hbm_gbs / bf16_tflops_sustained)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md 6.65 TB/s, ~1.4 PFLOP/s sustained)"


# ---------------------------------------------------------------------------------------------------------
# CPU oracle port (cpu_baseline, --impl reference)
# ---------------------------------------------------------------------------------------------------------
def cpu_models(a, vocab):
    """Oracle models of the benchmark shapes on the host.  Timing does not depend on weight values, so the
    decoder layers share one set of random tensors (keeps the 7B shape at ~2 GB instead of 27 GB)."""
    from oracle import llama_ref as LR
    out = []
    for name, seed in ((a.target, 1), (a.draft, 2)):
        s = SHAPES[name] if not corr_layers(name) else dict(SHAPES[a.target], n_layers=corr_layers(name))
        sh = LR.LlamaShape(vocab, s["hidden"], 1, s["n_heads"], s["mlp"])
        W = LR.make_weights(sh, seed, std=0.02)
        W["layers"] = W["layers"] * s["n_layers"]
        sh.n_layers = s["n_layers"]
        out.append(LR.RefLlama(sh, W, "fp32"))
    return out


def cpu_one_user(a, models, ds, fn, u):
    from oracle import bssd_ref
    t0 = time.perf_counter()
    res = bssd_ref.bssd(models[0], models[1], ds.prompt_ids(u), a.K, a.N, a.gamma, 4, fn)
    return time.perf_counter() - t0, res


def make_fn(ds, kind):
    from atspeed_b200.generation_trie import (Trie, positional_prefix_allowed_tokens_fn, suffix_prefix_allowed_tokens_fn)
    from atspeed_b200.prompts import RESPONSE_SEP
    if kind == "strict":
        return suffix_prefix_allowed_tokens_fn(Trie(ds.strict_trie_sequences()), RESPONSE_SEP)
    return positional_prefix_allowed_tokens_fn(ds.positional_allowed(), RESPONSE_SEP)


def host_threads():
    """All the host cores this process may use (torchrun pins OMP_NUM_THREADS=1: undo that for the CPU arm)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def reference_arm(a, rank):
    if rank != 0:
        return
    host_threads()
    from atspeed_b200.prompts import load_dataset
    ds = load_dataset(a.dataset)
    fn = make_fn(ds, a.constraint)
    models = cpu_models(a, ds.vocab_size)
    users = list(range(ds.n_users))
    for w in range(a.warmup):
        log("reference arm (CPU oracle port): warm-up user %d" % w)
        cpu_one_user(a, models, ds, fn, users[w % len(users)])
    t0 = time.perf_counter()
    for s in range(a.steps):
        log("reference arm (CPU oracle port): timed user %d" % s)
        cpu_one_user(a, models, ds, fn, users[(a.warmup + s) % len(users)])
    dt = time.perf_counter() - t0
    val = a.steps / dt
    cores = torch.get_num_threads()
    sample = "1 user per step (bounded sample of the per-step user batch); oracle port, fp32, layer weights shared across layers"
    print(json.dumps({"impl": "reference", "metric": "topk_recs_per_sec", "value": val, "unit": "users/s", "n_gpus": a.gpus,
                      "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": workload_name(a)},
                      "cpu_baseline": {"value": val, "unit": "users/s", "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------------------------------------
def atspeed_arm(a, rank, world, local_rank):
    import torch.distributed as dist
    from atspeed_b200 import _lib
    from atspeed_b200.constraint import compile_constraint
    from atspeed_b200.engine import DeviceModel, DeviceTrie, ModelSpec, Session
    from atspeed_b200.prompts import load_dataset
    from atspeed_b200.runner import shard_users

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    log("setup: models, sessions, prompts")
    ds = load_dataset(a.dataset)
    V = ds.vocab_size
    fn = make_fn(ds, a.constraint)
    specs = []
    for name in (a.target, a.draft):
        s = SHAPES[name] if not corr_layers(name) else dict(SHAPES[a.target], n_layers=corr_layers(name))
        specs.append(ModelSpec(V, s["hidden"], s["n_layers"], s["n_heads"], s["hidden"] // s["n_heads"], s["mlp"]))
    tdm = DeviceModel(specs[0], gpu_weights(specs[0], 1, dev), dev)
    if corr_layers(a.draft):
        ddm = DeviceModel(specs[1], correlated_draft_weights(tdm, corr_layers(a.draft), 0.03, dev), dev)
    else:
        ddm = DeviceModel(specs[1], gpu_weights(specs[1], 2, dev), dev)
    csr = compile_constraint(fn, ds.prompt_ids(0), 4, other_prompt=ds.prompt_ids(1))
    dtrie = DeviceTrie(csr, dev)
    skw = dict(do_sample=True, top_k=50, temperature=1.0, seed=2025) if a.do_sample else {}
    n_lanes = max(1, a.lanes)
    if a.cohort > 1:
        skw["max_users"] = a.cohort
        skw["cohort_tokens"] = a.cohort_tokens
    lanes = [Session(tdm, ddm, dtrie, a.K, a.N, 4, **skw) for _ in range(n_lanes)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)]
    sess = lanes[0]
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(n_lanes)
    U = a.users_per_step
    mine = shard_users(list(range(ds.n_users)), rank, world)
    if a.emulate_shard and world == 1:
        er, ew = (int(x) for x in a.emulate_shard.split("/"))
        mine = shard_users(list(range(ds.n_users)), er, ew)
    n_steps_total = a.warmup + a.steps
    step_users = [[mine[(s * U + i) % len(mine)] for i in range(U)] for s in range(n_steps_total)]
    prompts_host = {u: ds.prompt_ids(u) for us in step_users for u in us}
    prompts_dev = {u: torch.tensor(p, dtype=torch.int32, device=dev) for u, p in prompts_host.items()}
    # the opening tokens every prompt of this run shares (the instruction template): their K/V is computed once per session
    shared_prefix = 0
    if a.cohort > 1 and not a.no_shared_prefix:
        from atspeed_b200.runner import common_prefix
        pre = common_prefix(list(prompts_host.values()))
        if pre:
            for ss in lanes:
                shared_prefix = ss.set_shared_prefix(pre)
    tok_dev = torch.zeros(U, a.K, _lib.MAX_NEW, dtype=torch.int32, device=dev)
    sc_dev = torch.zeros(U, a.K, dtype=torch.float32, device=dev)
    gathered = [torch.zeros_like(tok_dev) for _ in range(world)] if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def on_lanes(fn, s):
        """Users of step s dealt round-robin to the lanes; every lane runs on its own stream from its own thread (the C
        calls release the GIL); the default stream is fenced before and after with events.  fn(session, lane, indices)
        handles the lane's users (indices into the step's user list) and returns a list."""
        start = torch.cuda.Event()
        start.record()

        def work(l):
            torch.cuda.set_device(dev)
            with torch.cuda.stream(streams[l]):
                streams[l].wait_event(start)
                out = fn(lanes[l], l, list(range(l, len(step_users[s]), n_lanes)))
                done = torch.cuda.Event()
                done.record()
            return out, done

        res = list(pool.map(work, range(n_lanes)))
        for _, done in res:
            torch.cuda.current_stream(dev).wait_event(done)
        return [x for out, _ in res for x in out]

    # cohort mode: each lane's users of a step as ONE concatenated device tensor (inputs resident in HBM)
    cat_dev = {}
    if a.cohort > 1:
        for s_ in range(n_steps_total):
            for l in range(n_lanes):
                us = [step_users[s_][i] for i in range(l, U, n_lanes)]
                cat_dev[(s_, l)] = (torch.cat([prompts_dev[u] for u in us]), [len(prompts_host[u]) for u in us])
    lane_tok = [torch.zeros((U + n_lanes - 1) // n_lanes, a.K, _lib.MAX_NEW, dtype=torch.int32, device=dev) for _ in range(n_lanes)]
    lane_sc = [torch.zeros((U + n_lanes - 1) // n_lanes, a.K, dtype=torch.float32, device=dev) for _ in range(n_lanes)]

    # device-resident pass: the inputs are the users' history item ids, resident in HBM; the prompts are built on the device
    # (csrc/prompt.cu) inside the timed region
    from atspeed_b200.prompts import DevicePromptBuilder
    builder = DevicePromptBuilder(ds, dev)

    def step_device(s, trace=False):
        if a.cohort > 1:
            def fn(ss, l, idx):
                cat, lens = builder.build([step_users[s][i] for i in idx])
                sts = ss.bssd_batch_device(cat, lens, a.gamma, lane_tok[l], lane_sc[l])
                tok_dev[idx] = lane_tok[l][: len(idx)]
                return sts
        else:
            def fn(ss, l, idx):
                return [ss.bssd_device(prompts_dev[step_users[s][i]], a.gamma, tok_dev[i], sc_dev[i]) for i in idx]
        sts = on_lanes(fn, s)
        if trace:
            log("  step %d: searches of this rank done" % s)
        if world > 1:   # the one collective of the path: ranked lists of every rank, over NVLink
            dist.all_gather(gathered, tok_dev)
            if trace:
                torch.cuda.synchronize(dev)
                log("  step %d: all-gather done" % s)
        return (sum(st["kernel_launches"] for st in sts), sum(st["total_accept_steps"] for st in sts),
                sum(st["n_run"] for st in sts))

    host_lists = {}

    def step_host(s):
        if a.cohort > 1:
            def fn(ss, l, idx):
                t0 = time.perf_counter()
                outs = ss.bssd_batch([prompts_host[step_users[s][i]] for i in idx], a.gamma)
                dt = time.perf_counter() - t0
                return [(dt, o["tokens"]) for o in outs]          # every user of the call waits for the whole cohort
        else:
            def fn(ss, l, idx):
                out = []
                for i in idx:
                    t0 = time.perf_counter()
                    o = ss.bssd(prompts_host[step_users[s][i]], a.gamma)
                    out.append((time.perf_counter() - t0, o["tokens"]))
                return out
        res = on_lanes(fn, s)
        # ranked lists of the step as one fixed-shape block (a user with fewer than K beams is padded with zeros)
        block = np.zeros((U, a.K, 4), dtype=np.int32)
        order = [i for l in range(n_lanes) for i in range(l, U, n_lanes)]       # on_lanes returns lane-major
        for i, (_, toks) in zip(order, res):
            block[i, : toks.shape[0], : toks.shape[1]] = toks[: a.K, :4]
        host_lists[s] = block
        if world > 1:
            t = torch.from_numpy(block).to(dev)
            dist.all_gather([torch.empty_like(t) for _ in range(world)], t)
        return [x[0] for x in res]

    # ---- device-resident pass (value) ----
    for s in range(a.warmup):
        log("device pass: warm-up step %d" % s)
        step_device(s, trace=True)
    log("device pass: barrier before the timed region")
    barrier()
    log("device pass: timed region (%d steps)" % a.steps)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = accept = runs = 0
    # clocks are reported for rank 0's GPU only: one nvidia-smi poller per box, not one per rank
    with ClockSampler(local_rank if rank == 0 else None) as clocks:
        ev0.record()
        for s in range(a.warmup, n_steps_total):
            l, ac, rn = step_device(s)
            launches += l; accept += ac; runs += rn
        ev1.record()
        barrier()
    dev_lists_last = tok_dev.cpu().numpy()[:, :, :4].copy()          # device-resident pass, last timed step
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    users_total = world * U * a.steps
    base = {"metric": "topk_recs_per_sec", "value": users_total / (ms_total * 1e-3), "unit": "users/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_total / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(a), "l2": "inputs larger than L2 (13.5 GB of weights streamed per forward)",
                       "inputs": "history item ids resident in HBM; prompts built on the device inside the timed region (cohort mode)",
                       "parallelism": f"user-sharded x{world}, one all-gather of ranked lists per step",
                       "gemm_pair_kernel": os.environ.get("ATSPEED_GEMM_2CTA", "1") != "0",
                       "shared_prompt_prefix_tokens": shared_prefix},
            "clocks": clocks.summary(), "gpu_launches": int(launches),
            "accepted_tokens_per_verify": accept * a.K / max(1, runs)}
    STATE["partial"] = dict(base)          # what the watchdog prints if a later phase stalls
    log("device pass done: %.1f users/s; host-buffer pass: warm-up" % base["value"])
    # ---- host-buffer pass (e2e) ----
    for s in range(min(a.warmup, 1)):
        step_host(s)
    barrier()
    log("host-buffer pass: timed region")
    t0 = time.perf_counter()
    lat = []
    for s in range(a.warmup, n_steps_total):
        lat += step_host(s)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    # integrity: the host-buffer pass re-runs the same users in the same cohorts, so its ranked lists must be the
    # device-resident pass's, bit for bit (a difference would mean a race between lanes / streams)
    try:
        if a.do_sample:         # every search draws its own noise stream: the two passes are different samples
            consistency = {"skipped": "sampling mode"}
        else:
            same = int((host_lists[n_steps_total - 1] == dev_lists_last).all(axis=(1, 2)).sum())
            consistency = {"users_compared": U, "identical_ranked_lists": same}
    except Exception as e:      # informative only
        consistency = {"error": repr(e)[:200]}
    h2d = int(np.mean([sum(len(prompts_host[u]) * 4 for u in step_users[s]) for s in range(a.warmup, n_steps_total)]))
    d2h = U * (a.K * 4 * 4 + a.K * 4) + int(round(runs / max(a.steps, 1))) * 64
    e2e = {"value": users_total / float(e2e_s.item()), "unit": "users/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
    STATE["partial"] = dict(base, e2e=e2e, pass_consistency=consistency)
    log("host-buffer pass done: %.1f users/s (%s); single-search latency" % (e2e["value"], consistency))
    # single-search latency (nothing else in flight on the GPU): what one user waits for; the ranked lists are kept for the
    # parity record against HF generate on the same weights
    lat1, ours = [], {}
    hf_users = step_users[a.warmup][: min(U, max(8, a.hf_baseline_users))]
    for u in hf_users:
        t0 = time.perf_counter()
        o = sess.bssd_batch([prompts_host[u]], a.gamma)[0] if a.cohort > 1 else sess.bssd(prompts_host[u], a.gamma)
        lat1.append(time.perf_counter() - t0)
        ours[u] = (o["tokens"], o["scores"])
    out = dict(base)
    out.update({"e2e": e2e,
                "latency_ms_p50": float(np.percentile(np.asarray(lat1) * 1e3, 50)),
                "latency_ms_p95": float(np.percentile(np.asarray(lat1) * 1e3, 95)),
                "latency_ms_p50_loaded": float(np.percentile(np.asarray(lat) * 1e3, 50)),
                "pass_consistency": consistency})
    STATE["partial"] = dict(out)
    if rank == 0 and world == 1 and a.check_users > 0 and not a.do_sample:
        out["parity_vs_oracle"] = check_users_against_oracle(a, ds, fn, tdm, ddm, sess, hf_users[: a.check_users], prompts_host)
        STATE["partial"] = dict(out)
    # ---- CPU baseline (rank 0, N = 1): the oracle port on the host cores, bounded sample ----
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            log("cpu_baseline: users through the oracle port")
            host_threads()
            models = cpu_models(a, V)
            dt, n_cpu = 0.0, 0
            while dt < 10.0 and n_cpu < 8:           # bounded sample: >= 10 s of CPU work or 8 users, whichever comes first
                t1, _ = cpu_one_user(a, models, ds, fn, step_users[a.warmup][n_cpu % U])
                dt, n_cpu = dt + t1, n_cpu + 1
                log("cpu_baseline: %d user(s), %.1f s" % (n_cpu, dt))
            del models
            out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "users/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": "%d user(s) of the same workload through oracle/bssd_ref.py (fp32, layer weights "
                                             "shared across layers); %.1f s of CPU work" % (n_cpu, dt)}
        else:
            out["cpu_baseline"] = None
        STATE["partial"] = dict(out)
    # ---- profiled pass (roofline of the dominant kernel, share of step per kernel group) ----
    roofline, groups = None, None
    if rank == 0:
        log("profiled pass (per-launch CUDA events)")
        sess.profile(True)
        if a.cohort > 1:
            a.profile_users = len(cat_dev[(a.warmup, 0)][1])
            sess.bssd_batch_device(*cat_dev[(a.warmup, 0)], a.gamma, lane_tok[0], lane_sc[0])
        else:
            for u in step_users[a.warmup][: max(1, a.profile_users)]:
                sess.bssd_device(prompts_dev[u], a.gamma, tok_dev[0], sc_dev[0])
        prof = sess.profile_read()
        sess.profile(False)
        tot = sum(v["ms"] for v in prof.values())
        groups = {k: {"ms_per_user": v["ms"] / max(1, a.profile_users), "launches_per_user": v["launches"] / max(1, a.profile_users),
                      "share": v["ms"] / tot if tot else 0.0,
                      # algorithmic bytes / CUDA-event time where the runtime knows the bytes on the host (GEMMs, kernel (a):
                      # rows x V x 4 B of fp32 logits -- L2-resident right after lm_head, so not an HBM figure)
                      "algorithmic_gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["bytes"] > 0 and v["ms"] > 0 else None}
                  for k, v in prof.items()}
        g = prof["gemm"]
        peak, peak_tf, how = peaks()
        sec = g["ms"] * 1e-3
        ach = g["bytes"] / sec / 1e9 if sec else 0.0
        ach_tf = g["flops"] / sec / 1e12 if sec else 0.0
        # which resource bounds the launches in aggregate: time the algorithmic bytes need at the measured copy peak vs
        # time the FLOPs need at the measured sustained cuBLAS peak (cohort forwards are large enough to be tensor-bound)
        t_hbm, t_tensor = g["bytes"] / (peak * 1e9), g["flops"] / (peak_tf * 1e12)
        kname = "gemm_wx_tcgen05" + (("_2cta<%s>" % ("4" if os.environ.get("ATSPEED_GEMM_CLUSTER") == "4" else "2"))
                                     if base["config"]["gemm_pair_kernel"] else "")
        traffic, tinfo = ncu_traffic(kname, "cohort" if a.cohort > 1 else "single")
        if t_tensor > t_hbm:
            roofline = {"kernel": kname, "bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": ach_tf / peak_tf, "hbm_achieved_gbs": ach, "hbm_frac": ach / peak}
        else:
            roofline = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "tensor_achieved_tflops": ach_tf, "tensor_frac": ach_tf / peak_tf}
        roofline.update({"traffic": traffic, "traffic_source": tinfo, "peak_source": how,
                         "avg_launch_us": g["ms"] * 1e3 / max(1, g["launches"]), "launches": g["launches"],
                         "algorithmic_bytes_per_launch": g["bytes"] / max(1, g["launches"]),
                         "flops_per_launch": g["flops"] / max(1, g["launches"]),
                         "roofline_time_frac": max(t_hbm, t_tensor) / sec if sec else 0.0})
    out.update({"kernel_groups": groups, "roofline": roofline})
    STATE["partial"] = dict(out)
    if rank == 0:
        # ---- HF generate on the same GPU, same shape, same weights: the north star's comparison (N = 1 only) ----
        if world == 1 and a.hf_baseline_users > 0:
            log("hf_gpu_baseline: transformers generate(num_beams=K) on the same GPU")
            box = {}

            def run_hf():
                try:
                    W = {"embed": tdm.embed, "norm": tdm.norm, "lm_head": tdm.lm_head, "layers": tdm.layers}
                    box["r"] = hf_baseline(a, ds, fn, dev, hf_users[: a.hf_baseline_users], W, ours, sess)
                except Exception as e:   # informative only: never fail the bench line on the comparison arm
                    box["r"] = {"error": repr(e)[:300]}

            th = threading.Thread(target=run_hf, daemon=True)
            th.start()
            th.join(float(os.environ.get("ATSPEED_BENCH_HF_LIMIT_S", 240)))
            hf = box.get("r") or {"error": "timed out"}
            if "users_per_s" in hf:
                hf["speedup_e2e_throughput"] = out["e2e"]["value"] / hf["users_per_s"]           # cohorts x lanes vs HF
                hf["speedup_single_search"] = hf["latency_ms_p50"] / out["latency_ms_p50"]       # one user alone vs HF (speedupTF)
            out["hf_gpu_baseline"] = hf
        STATE["partial"] = None
        print(json.dumps(out), flush=True)
        STATE["printed"] = True
    if world > 1:
        # nobody tears its communicator down while rank 0 is still in its profiled pass
        log("final barrier")
        barrier()
    log("done")


def hf_baseline(a, ds, fn, dev, users, weights=None, ours=None, sess=None):
    """HF `generate(num_beams=K, prefix_allowed_tokens_fn=...)` on the same GPU, shape and -- when `weights` is given -- the
    same random-init weights (reference code/inference.py:177-178, the `TF_target` column): the north star's >= 2x
    comparison, one user at a time as the reference runs it.  3 untimed warm-up calls, then every user of `users`.
    `ours`: {user: (tokens [K,4], scores [K])} from this repo's single-search runs of the same users -> a parity record at
    the benchmark shape (ranked lists identical / explained by a bf16 near-tie; HF's sequences_scores x 4 = summed
    log-probs).  Never fails the bench line."""
    from transformers import LlamaConfig, LlamaForCausalLM
    s = SHAPES[a.target]
    cfg = LlamaConfig(vocab_size=ds.vocab_size, hidden_size=s["hidden"], intermediate_size=s["mlp"],
                      num_hidden_layers=s["n_layers"], num_attention_heads=s["n_heads"], num_key_value_heads=s["n_heads"],
                      tie_word_embeddings=False, pad_token_id=0, bos_token_id=1, eos_token_id=2)
    with torch.device(dev):
        m = LlamaForCausalLM(cfg)
    # `.to(bfloat16)` would also round the rotary embedding's inv_freq buffer, which from_pretrained(dtype=bf16) keeps in fp32
    # (positions x a bf16-rounded frequency are off by up to ~0.2 rad at position 100: measured as a 0.05 systematic logit
    # offset against a true fp32 forward, tools/forward_accuracy.py): convert, then restore the buffer
    inv_freq = m.model.rotary_emb.inv_freq.detach().clone().float()
    m = m.to(torch.bfloat16).eval()
    m.model.rotary_emb.inv_freq = inv_freq
    if hasattr(m.model.rotary_emb, "original_inv_freq"):
        m.model.rotary_emb.original_inv_freq = inv_freq
    same_weights = False
    if weights is not None:
        sd = {"model.embed_tokens.weight": weights["embed"], "model.norm.weight": weights["norm"], "lm_head.weight": weights["lm_head"]}
        names = {"wq": "self_attn.q_proj", "wk": "self_attn.k_proj", "wv": "self_attn.v_proj", "wo": "self_attn.o_proj",
                 "wg": "mlp.gate_proj", "wu": "mlp.up_proj", "wd": "mlp.down_proj", "ln1": "input_layernorm",
                 "ln2": "post_attention_layernorm"}
        for i, ly in enumerate(weights["layers"]):
            for k, n in names.items():
                sd[f"model.layers.{i}.{n}.weight"] = ly[k]
        missing, unexpected = m.load_state_dict(sd, strict=False)
        same_weights = not [k for k in missing if "rotary" not in k and "inv_freq" not in k]
    logit_cmp, cmp_rows = None, []
    if sess is not None and same_weights:
        # the forward itself at the benchmark shape: last-position logits of the prompt from this repo's kernels and from the
        # HF bf16 module (same weights); below, after the timed calls, the SAME HF module in fp32 is the yardstick for both
        n_cmp = min(3, len(users))
        for u in users[:n_cmp]:
            prompt = ds.prompt_ids(u)
            P = len(prompt)
            i32 = lambda x: torch.tensor(list(x), dtype=torch.int32, device=dev)
            mine = sess.forward_raw(0, i32(prompt), i32(range(P)), i32(range(P)), i32(range(1, P + 1)),
                                    torch.zeros(P, 16, dtype=torch.int32, device=dev), P, P, i32([P - 1]))[0]
            with torch.no_grad():
                theirs = m(input_ids=torch.tensor([prompt], device=dev)).logits[0, -1].float().cpu().numpy()
            cmp_rows.append((prompt, mine, theirs))
    lat, lists = [], {}
    seq = [users[0]] * 3 + list(users)                              # three untimed warm-up calls (lazy init, autotuning)
    for i, u in enumerate(seq):
        ids = torch.tensor([ds.prompt_ids(u)], device=dev)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        with torch.no_grad():
            o = m.generate(input_ids=ids, num_beams=a.K, num_return_sequences=a.K, max_new_tokens=4, do_sample=False,
                           prefix_allowed_tokens_fn=fn, use_cache=True, pad_token_id=0, output_scores=True,
                           return_dict_in_generate=True, length_penalty=1.0, early_stopping=False)
        torch.cuda.synchronize(dev)
        if i >= 3:
            lat.append(time.perf_counter() - t0)
            sc = getattr(o, "sequences_scores", None)
            lists[u] = (o.sequences[:, ids.shape[1]:].cpu().numpy(), None if sc is None else (sc.float().cpu().numpy() * 4.0))
    if cmp_rows:
        try:
            m = m.float()                                   # the same module and weights in fp32: the yardstick
            m.model.rotary_emb.inv_freq = inv_freq
            lo, hi = ds.level_ranges()[0]
            acc = {"ours_vs_hf_bf16": [], "ours_vs_hf_fp32": [], "hf_bf16_vs_hf_fp32": [], "std": [], "ov_ours": [], "ov_hf": []}
            for prompt, mine, theirs in cmp_rows:
                with torch.no_grad():
                    ref = m(input_ids=torch.tensor([prompt], device=dev)).logits[0, -1].float().cpu().numpy()
                acc["ours_vs_hf_bf16"].append(float(np.abs(mine - theirs).mean()))
                acc["ours_vs_hf_fp32"].append(float(np.abs(mine - ref).mean()))
                acc["hf_bf16_vs_hf_fp32"].append(float(np.abs(theirs - ref).mean()))
                acc["std"].append(float(ref.std()))
                top = lambda x: set(np.argsort(-x[lo:hi + 1])[:10])
                acc["ov_ours"].append(len(top(mine) & top(ref)) / 10.0)
                acc["ov_hf"].append(len(top(theirs) & top(ref)) / 10.0)
            logit_cmp = {"prompts": len(cmp_rows), "logit_std": float(np.mean(acc["std"])),
                         "mean_abs_diff_ours_vs_hf_fp32": float(np.mean(acc["ours_vs_hf_fp32"])),
                         "mean_abs_diff_hf_bf16_vs_hf_fp32": float(np.mean(acc["hf_bf16_vs_hf_fp32"])),
                         "mean_abs_diff_ours_vs_hf_bf16": float(np.mean(acc["ours_vs_hf_bf16"])),
                         "top10_first_code_token_overlap_with_fp32": {"ours": float(np.mean(acc["ov_ours"])), "hf_bf16": float(np.mean(acc["ov_hf"]))}}
        except Exception as e:
            logit_cmp = {"error": repr(e)[:200]}
    del m
    torch.cuda.empty_cache()
    ms = np.asarray(lat) * 1e3
    out = {"users_per_s": len(lat) / float(sum(lat)), "latency_ms_p50": float(np.percentile(ms, 50)),
           "latency_ms_mean": float(ms.mean()), "latency_ms_min": float(ms.min()), "latency_ms_max": float(ms.max()),
           "users": len(lat), "same_weights_as_target": bool(same_weights), "prompt_logits_vs_ours": logit_cmp,
           "what": "transformers LlamaForCausalLM.generate(num_beams=K, prefix_allowed_tokens_fn) bf16, same GPU, same shape, "
                   "one user at a time as code/inference.py:177-178 runs it; users_per_s = users / sum of latencies"}
    if ours and same_weights:
        # parity record at the benchmark shape.  HF's bf16 modules and this repo's kernels round at the same points but sum
        # in different orders, and with random-init weights the beam margins are tiny, so rank-for-rank equality of ten
        # beams is rare; what must hold is that both searches find (almost) the same items with (almost) the same scores.
        tol = 6e-2                      # bf16 tolerance on cumulative log-probs (tests/_common.py BF16_SCORE_TOL)
        ident = top1 = n = 0
        overlap, dscore = [], []
        for u, (hf_t, hf_s) in lists.items():
            if u not in ours:
                continue
            t, sc = ours[u]
            a_l, b_l = [tuple(r) for r in t[:, :4].tolist()], [tuple(r) for r in hf_t[:, :4].tolist()]
            n += 1
            ident += int(a_l == b_l)
            top1 += int(a_l[:1] == b_l[:1])
            overlap.append(len(set(a_l) & set(b_l)) / max(1, len(b_l)))
            if hf_s is not None:
                sb = dict(zip(b_l, hf_s))
                dscore += [abs(float(x) - float(sb[k])) for k, x in zip(a_l, sc) if k in sb]
        out["parity_vs_ours"] = {"users": n, "identical_ranked_lists": ident, "same_top1": top1,
                                 "mean_item_overlap": float(np.mean(overlap)) if overlap else None,
                                 "min_item_overlap": float(np.min(overlap)) if overlap else None,
                                 "max_abs_score_diff_common_items": float(np.max(dscore)) if dscore else None,
                                 "mean_abs_score_diff_common_items": float(np.mean(dscore)) if dscore else None,
                                 "tolerance": tol,
                                 "note": "informative: the rigorous check at this shape is --check-users (oracle, same bf16 "
                                         "contract) and tests/test_zz_gpu_shape7b.py"}
    return out


def arm_watchdog(total_s, stall_s):
    """A stalled GPU or collective must not hang the caller.  When the run exceeds `total_s`, or no phase breadcrumb has been
    written for `stall_s`, every thread's Python stack goes to stderr (where did the host stop?), rank 0 prints what it has
    measured so far -- marked incomplete -- and the process exits with code 17 (torchrun then stops the other ranks).  The
    only os._exit(0) is the case where the complete result line is already out and only the teardown stalled."""
    def fire(why):
        import faulthandler
        sys.stderr.write("bench.py watchdog [r%d]: %s; last phase: %s\n" % (STATE["rank"], why, STATE["phase"]))
        try:
            faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        except Exception:
            pass
        sys.stderr.flush()
        if STATE.get("printed"):      # the result line is out; only the teardown stalled
            os._exit(0)
        part = STATE.get("partial")
        if part is not None and STATE["rank"] == 0:      # whatever was measured before the stall, marked as such
            part["incomplete"] = "stalled in phase '%s'; keys measured after it are absent" % STATE["phase"]
            print(json.dumps(part), flush=True)
        os._exit(17)              # a stalled run is a FAILED run, with or without a partial line

    def watch():
        while True:
            time.sleep(2.0)
            now = time.perf_counter()
            if now - T_START > total_s:
                fire("no result after %d s" % total_s)
            if now - STATE["t_phase"] > stall_s:
                fire("no progress for %d s" % stall_s)

    t = threading.Thread(target=watch, daemon=True)
    t.start()
    return t


def multi_gpu_env(world):
    """Communicator settings for N > 1 (only defaults: anything the caller exports wins): the one collective moves ~12 KB
    per step, so NVSwitch multicast (NVLS) buys nothing -- leave its set-up out of the communicator's initialisation.  The
    kernels are the same at every N (config.gemm_pair_kernel)."""
    if world > 1:
        os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
        os.environ.setdefault("NCCL_MNNVL_ENABLE", "0")      # one box: no multi-node NVLink / IMEX probing either
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the one JSON line


def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    gpu_multi = world > 1 and a.impl != "reference"
    gpu_single = world == 1 and a.impl != "reference"
    arm_watchdog(int(os.environ.get("ATSPEED_BENCH_WATCHDOG_S", 600 if gpu_multi else 900)),
                 int(os.environ.get("ATSPEED_BENCH_STALL_S", 240 if gpu_multi else (300 if gpu_single else 600))))
    if a.impl == "reference":
        reference_arm(a, rank)
        return
    multi_gpu_env(world)
    try:
        if world > 1:
            import datetime
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            log("init_process_group(nccl), world %d" % world)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=900))
            torch.cuda.set_device(local_rank)
            log("communicator up; first collective (barrier)")
            dist.barrier()
            torch.cuda.synchronize()
            log("barrier done")
        atspeed_arm(a, rank, world, local_rank)
    except BaseException:
        # fail fast and loudly: a rank that raises must not sit in destroy_process_group while its peers wait in a collective
        import traceback
        sys.stderr.write("bench.py [r%d] failed in phase '%s':\n%s" % (rank, STATE["phase"], traceback.format_exc()))
        sys.stderr.flush()
        sys.stdout.flush()
        os._exit(1)
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        # every rank has passed its last collective and rank 0 has printed: tear the communicator down with a bound (the
        # teardown can wait on peers), then leave through the interpreter's normal exit so atexit hooks run
        import torch.distributed as dist
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(30.0)
        if t.is_alive():
            sys.stderr.write("bench.py [r%d]: communicator teardown did not finish in 30 s\n" % rank)
            os._exit(0)           # the result line is out; only the teardown stalled
    STATE["t_phase"] = time.perf_counter()
    return 0


if __name__ == "__main__":
    sys.exit(main())
