#!/usr/bin/env python
"""bench.py -- VMLMF-LSTM training / inference throughput (sequences/s) on B200, with per-kernel rooflines and the
same-run host-CPU baseline.

    python bench.py [--config cfg2] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads = BASELINE.json `configs` (synthetic data of the named shapes, weights from the reference initialisers under
torch.manual_seed(3); datasets are not available offline):
  cfg1  Net(9,[128],w_rank=8,u_rank=[6])            x[64,128,9]     UCI-HAR shape at the reference's batch
  cfg2  Net(77,[256],w_rank=8,u_rank=[6])           x[B,24,77]      Opportunity windows        <- headline (default)
  cfg3  Net(9,[128],8,[2,4],cell=MyVMLMFCellg2)     x[8192,128,9]   group-structured cell
  cfg4  Model(10000,650,2,.,0.05,300,[300],"vmlmf") tok[35,B]       PTB-shaped language model, B=512 per GPU
  cfg5  Net(9,[1024],w_rank=64,u_rank=[64])         x[2048,128,9]   scaling sweep point (cfg5b: H=4096, rank 256)
One "step" is the reference's training iteration: HAR = zero_grad, forward, cross-entropy, backward, Adam(lr 0.002)
(V/train_test/train.py:58-65); LM = detach carried state, forward, nll_loss, backward, clip_grad_norm_(5) + SGD
(V/train_test/lm_test.py:196-209).  With N>1 the batch is sharded over ranks (fixed per-GPU batch => weak scaling) and
the live gradients are reduced with ONE flat-bucket NCCL all-reduce per step, captured inside the step's CUDA graph
together with the optimizer (HAR: mean over ranks = batch-mean loss; LM: sum over ranks = token-mean x GLOBAL batch,
clip on the post-allreduce norm).

One JSON line on stdout (rank 0).  `value` = whole-job train sequences/s of `--config` with inputs resident in HBM;
`e2e` = the same step fed from PINNED HOST batches through vmlmf_b200.data.SyntheticLoader (H2D copy of every batch
and a D2H read of every loss inside the timed region); `roofline` = the dominant C-ABI call timed live with CUDA
events; `configs` = the same measurement (value, inference, roofline) for every BASELINE config (default run, N=1);
`cpu_baseline` = the CPU oracle port of the reference (oracle/vmlmf_oracle.py, torch eager, all host threads) on a
bounded sample of the headline workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vmlmf_lstm_train_sequences_per_sec"

CONFIGS = {
    "cfg1": dict(kind="har", I=9, H=128, wr=8, ur=[6], T=128, classes=6, batch=64, cell="plain",
                 desc="cfg1: VMLMF LSTM Net(9,[128],w_rank=8,u_rank=[6]) on synthetic UCI-HAR windows [64,128,9], 6 classes, train step (fwd+CE+bwd+Adam)"),
    "cfg2": dict(kind="har", I=77, H=256, wr=8, ur=[6], T=24, classes=18, batch=9472, cell="plain",
                 desc="cfg2: VMLMF LSTM Net(77,[256],w_rank=8,u_rank=[6]) on synthetic Opportunity windows [B,24,77], 18 classes, train step (fwd+CE+bwd+Adam)"),
    "cfg3": dict(kind="har", I=9, H=128, wr=8, ur=[2, 4], T=128, classes=6, batch=8192, cell="group",
                 desc="cfg3: group-structured VMLMF cell Net(9,[128],8,[2,4],cell=MyVMLMFCellg2) on HAR windows [8192,128,9], train step"),
    "cfg4": dict(kind="lm", V=10000, H=650, layers=2, wr=300, ur=[300], T=35, batch=512, dropout=0.5,
                 desc="cfg4: vmlmf_lm Model(10000,650,2,0.5,0.05,300,[300],'vmlmf') on PTB-shaped synthetic tokens [35,B], B=512 per GPU, train step (fwd+nll_loss+bwd+clip 5+SGD), carried state"),
    "cfg4_b20": dict(kind="lm", V=10000, H=650, layers=2, wr=300, ur=[300], T=35, batch=20, dropout=0.5,
                     desc="cfg4 at the reference's own batch: Model(10000,650,2,0.5,0.05,300,[300],'vmlmf'), tokens [35,20] (lm_test.py defaults), train step"),
    "cfg5": dict(kind="har", I=9, H=1024, wr=64, ur=[64], T=128, classes=6, batch=2048, cell="plain",
                 desc="cfg5: scaling-sweep point Net(9,[1024],w_rank=64,u_rank=[64]) on windows [2048,128,9], train step"),
    "cfg5b": dict(kind="har", I=9, H=4096, wr=256, ur=[256], T=128, classes=6, batch=1024, cell="plain",
                  desc="cfg5b: scaling-sweep point Net(9,[4096],w_rank=256,u_rank=[256]) on windows [1024,128,9], train step"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    # cfg2 default 9472 = 4 x 148 SMs x 16: the warp-MMA kernels tile 16 sequences per CTA, so this batch is whole waves
    ap.add_argument("--batch", type=int, default=0, help="sequences per GPU per step (0 = the config's default)")
    ap.add_argument("--cpu-batch", type=int, default=0, help="sequences per CPU-baseline step (0 = bounded default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config array (default run, N=1 only)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of the step eagerly instead of replaying a CUDA graph")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- #
# algorithmic work per (sequence, timestep) -- SURVEY.md 8(d)
# ----------------------------------------------------------------------------------------------- #

def algorithmic(c):
    """bytes and flops per unit (one sequence-timestep, all layers) of a config"""
    H, wr, ur = c["H"], c["wr"], sum(c["ur"])
    layers = c.get("layers", 1)
    I = c["I"] if c["kind"] == "har" else H
    f_step = 2 * I * wr + 2 * wr * 4 * H + 8 * I + 2 * H * ur + 2 * ur * 4 * H + 8 * H + 17 * H
    return dict(q_inf=4 * (I + H) * layers, q_fwd=4 * (I + 6 * H) * layers, q_bwd=4 * (2 * I + 7 * H) * layers,
                q_train=4 * (3 * I + 13 * H) * layers, f_step=f_step * layers)


def peaks():
    p = {}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    src = "measured (MEASURED_PEAKS.json)" if p else "fallback (B200_PROFILING.md)"
    hbm = float(p.get("hbm_gbs", 6650.0))
    bf16 = float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0)))
    return hbm, bf16 / 2.0, src          # HBM GB/s; dense TF32 TFLOP/s = half the measured bf16 rate (inside a long step: sustained)


def roofline_of(name, ms, units, q_bytes, flops, regime, traffic=None):
    """the binding roofline of one call: HBM (algorithmic bytes) or tensor (algorithmic flops; the 3xTF32 products issue 3x)"""
    hbm, tf32, src = peaks()
    if not ms:
        return None
    t_hbm = q_bytes * units / (hbm * 1e9)
    t_tc = flops * units / (tf32 * 1e12)
    tensor_bound = regime == "R2" and t_tc > t_hbm
    if tensor_bound:
        ach = flops * units / (ms * 1e-3) / 1e12
        return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": tf32, "unit": "TFLOP/s", "frac": ach / tf32,
                "peak_source": src + ", dense TF32 = bf16 / 2", "ms_per_launch": ms, "algorithmic_flops_per_launch": flops * units,
                "note": "fp32-accurate products issue 3 TF32 MMAs each: the ceiling of this mode is frac = 1/3", "traffic": traffic}
    ach = q_bytes * units / (ms * 1e-3) / 1e9
    return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
            "peak_source": src, "ms_per_launch": ms, "algorithmic_bytes_per_launch": q_bytes * units, "traffic": traffic}


# ----------------------------------------------------------------------------------------------- #
# CPU arm: the oracle port of the reference, timed on the host cores
# ----------------------------------------------------------------------------------------------- #

def cpu_train_rate(cfg_name, batch, steps, warmup, budget_s=60.0):
    """sequences/s of the reference training iteration on CPU (oracle port, torch eager, all threads)."""
    import torch

    import vmlmf_b200 as vb
    from oracle import vmlmf_oracle as vo

    c = CONFIGS[cfg_name]
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(3)
    g = torch.Generator().manual_seed(1234)
    if c["kind"] == "har":
        cell_cls = vb.MyVMLMFCellg2 if c["cell"] == "group" else vb.MyVMLMFCell
        net = vb.Net(c["I"], [c["H"]], w_rank=c["wr"], u_rank=c["ur"], cell=cell_cls)        # parameters only (CPU)
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items() if not k.startswith("cell.")}
        pre = "rnn.rnncells.0." + ("layers." if c["cell"] == "group" else "")
        cell = vo.split_state_dict(sd, pre)
        opt = torch.optim.Adam(list(sd.values()), lr=0.002)
        x = torch.randn(batch, c["T"], c["I"], generator=g)
        y = torch.randint(0, c["classes"], (batch,), generator=g)

        def step():
            opt.zero_grad(set_to_none=True)
            logits = vo.net_forward([cell], sd["lin.weight"], sd["lin.bias"], x, kind=c["cell"])
            loss = torch.nn.functional.cross_entropy(logits, y)
            loss.backward()
            opt.step()
            return loss.item()
    else:
        m = vb.Model(c["V"], c["H"], c["layers"], 0.0, 0.05, w_rank=c["wr"], u_ranks=c["ur"], lstm_type="vmlmf")
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
        layers = [vo.split_state_dict(sd, f"rnns.{i}.") for i in range(c["layers"])]
        tok = torch.randint(0, c["V"], (c["T"], batch), generator=g)
        y = torch.randint(0, c["V"], (c["T"], batch), generator=g)
        states = [[torch.zeros(batch, c["H"]), torch.zeros(batch, c["H"])] for _ in range(c["layers"])]

        def step():
            for p in sd.values():
                p.grad = None
            st = [(h.detach(), cc.detach()) for h, cc in states]
            scores, new = vo.lm_model_forward(sd["embed.w"], layers, sd["fc.w"], sd["fc.b"], tok, st, drop_p=c["dropout"], training=True)
            loss = vo.lm_nll_loss(scores, y)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(list(sd.values()), 5.0)
            with torch.no_grad():
                for p in sd.values():
                    p -= 1.0 * p.grad
            for i, (h, cc) in enumerate(new):
                states[i] = [h, cc]
            return loss.item()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done * batch / dt, dt / done * 1e3, done, torch.get_num_threads()


CPU_BATCH = {"cfg1": 64, "cfg2": 1024, "cfg3": 512, "cfg4": 20, "cfg4_b20": 20, "cfg5": 64, "cfg5b": 16}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the reference is pure Python and does not travel to the
    GPU box), all host threads, same metric / config, a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b = args.cpu_batch or CPU_BATCH[args.config]
    rate, ms, done, threads = cpu_train_rate(args.config, b, args.steps, max(1, min(args.warmup, 3)), budget_s=150.0)
    sample = f"{done} steps x {b} sequences of the {args.config} train step on CPU (oracle port of the reference, torch eager)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "sequences/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CONFIGS[args.config]["desc"], "per_step_batch": b},
        "cpu_baseline": {"value": rate, "unit": "sequences/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------- #
# GPU arm
# ----------------------------------------------------------------------------------------------- #

class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # idle samples (before the first kernel ramps the clock) sit at the bottom; report the median
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_max_seen": max(sm) if sm else None, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index):
    """pin this rank's CPU affinity to the NUMA node its GPU hangs off, so that the pinned host batches it allocates next are
    node-local (8 ranks x 70 MB per 1.3 ms step otherwise all stream from whichever node the launcher started on)"""
    try:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = vis.split(",")[index] if vis else str(index)
        bus = subprocess.run(["nvidia-smi", "-i", phys, "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if len(bus.split(":")[0]) == 8:               # nvidia-smi prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"node": node, "cpus": len(allowed)}
    except Exception as e:                            # best effort: containers may hide sysfs
        sys.stderr.write(f"bench.py: NUMA binding skipped ({type(e).__name__}: {e})\n")
    return None


class Workload:
    """model + optimizer + one training step of a BASELINE config on this rank's GPU"""

    def __init__(self, name, dev, batch, world, rank):
        import torch

        import vmlmf_b200 as vb
        from vmlmf_b200.parallel import GradBucket, broadcast_parameters
        self.name, self.c, self.dev, self.B, self.world = name, CONFIGS[name], dev, batch, world
        c = self.c
        torch.manual_seed(3)
        if c["kind"] == "har":
            cell_cls = vb.MyVMLMFCellg2 if c["cell"] == "group" else vb.MyVMLMFCell
            self.net = vb.Net(c["I"], [c["H"]], w_rank=c["wr"], u_rank=c["ur"], cell=cell_cls).to(dev)
            self.bucket = GradBucket(self.net, average=True, symmetric=world > 1)   # batch-MEAN loss: average over ranks; peer-mapped bucket
            self.opt = vb.FlatAdam(self.bucket, lr=0.002)
            self.loss_fn = vb.cross_entropy
            self.states = None
        else:
            self.net = vb.Model(c["V"], c["H"], c["layers"], c["dropout"], 0.05, w_rank=c["wr"], u_ranks=c["ur"], lstm_type="vmlmf").to(dev)
            # LM loss = token-mean x batch (lm_test.py:147,153): the local loss already carries the local batch factor, so
            # the SUM over ranks is the gradient of the global-batch loss; clip on the post-allreduce norm (lm_test.py:204)
            self.bucket = GradBucket(self.net, average=False)
            self.opt = vb.FlatClipSGD(self.bucket, lr=1.0, max_norm=5.0)
            self.loss_fn = vb.nll_loss
            self.states = self.net.state_init(batch)
        broadcast_parameters(self.net)
        self.vb = vb

    def make_batch(self, g):
        import torch
        c = self.c
        if c["kind"] == "har":
            return torch.randn(self.B, c["T"], c["I"], generator=g), torch.randint(0, c["classes"], (self.B,), generator=g)
        return torch.randint(0, c["V"], (c["T"], self.B), generator=g), torch.randint(0, c["V"], (c["T"], self.B), generator=g)

    def forward_loss(self, x, y):
        if self.states is None:
            return self.loss_fn(self.net(x), y)
        st = self.net.detach(self.states)                       # truncated BPTT: carried, detached state (lm_test.py:199)
        scores, self._new_states = self.net(x, st)
        return self.loss_fn(scores, y)

    def carry_state(self):
        """keep the carried state in stable buffers (graph replays); after backward, which still reads the old values"""
        if self.states is not None:
            for (h, cc), (hn, cn) in zip(self.states, self._new_states):
                h.copy_(hn.detach()); cc.copy_(cn.detach())
            self._new_states = None

    def eager_step(self, x, y):
        self.bucket.zero()
        loss = self.forward_loss(x, y)
        loss.backward()
        self.carry_state()
        self.bucket.all_reduce()
        self.opt.step()
        return loss.detach()

    def infer(self, x):
        import torch
        with torch.no_grad():
            if self.states is None:
                return self.net(x)
            return self.net(x, self.net.detach(self.states))[0]

    def regime(self):
        from vmlmf_b200 import _lib
        c = self.c
        layer = self.net.rnn.rnncells[0] if c["kind"] == "har" else self.net.rnns[0]
        Ux, _, _, A = layer.canonical()[:4]                      # canonical factors: the group cell packs its blocks densely
        p = _lib.plan(c["T"], self.B, Ux.shape[0], c["H"], Ux.shape[1], A.shape[1]).path
        return {1: "R1", 2: "G", 3: "R1M", 4: "R2", 5: "R3"}.get(p, str(p))


def measure(w, K, W, use_graph, want_e2e, rank, world, dist):
    """train (resident inputs), optional e2e (pinned host inputs), inference; per-call kernel times from eager steps"""
    import torch

    from vmlmf_b200 import functional as F
    from vmlmf_b200.data import SyntheticLoader
    from vmlmf_b200.graphs import GraphedCallable
    dev = w.dev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return v

    resident = SyntheticLoader(w.make_batch, dev, pool=4, source="device", seed=1234 + rank)
    pool = resident.staging_buffers()
    for i in range(W):
        w.eager_step(*pool[i % len(pool)])
    F.EVENT_LOG = []
    n_log = 5
    for i in range(n_log):
        w.eager_step(*pool[i % len(pool)])
    torch.cuda.synchronize()
    log, F.EVENT_LOG = F.EVENT_LOG, None
    kt = {}
    for name, a, b in log:
        kt.setdefault(name, []).append(a.elapsed_time(b))
    kernel_ms = {k: sum(v) / n_log for k, v in kt.items()}       # per STEP (all layers)

    # One CUDA graph per input buffer: zero, forward, loss, backward, gradient pack, NCCL all-reduce (N > 1) and the
    # optimizer step all replay as one graph; the buffer a step reads is baked in, so no device-to-device input copy.
    graphs = {}
    graph_ok = {"collective_in_graph": world > 1 and use_graph}

    def build_graph(x, y):
        def whole():
            return w.eager_step(x, y)
        try:
            return GraphedCallable(whole, warmup=2)
        except Exception as e:                                     # a collective that cannot be captured on this stack
            if world == 1:
                raise
            sys.stderr.write(f"bench.py: capturing the all-reduce failed ({type(e).__name__}: {e}); keeping it eager\n")
            torch.cuda.synchronize()
            graph_ok["collective_in_graph"] = False

            def fwd_bwd():
                w.bucket.zero()
                loss = w.forward_loss(x, y)
                loss.backward()
                w.carry_state()
                w.bucket.pack()
                return loss.detach()
            fb = GraphedCallable(fwd_bwd, warmup=2)

            def step():
                loss = fb()
                w.bucket.all_reduce()
                w.opt.step()
                return loss
            return step

    def train_step(x, y):
        if not use_graph:
            return w.eager_step(x, y)
        g_ = graphs.get(x.data_ptr())
        if g_ is None:
            g_ = graphs[x.data_ptr()] = build_graph(x, y)
        return g_()

    for i in range(len(pool) + 3):
        train_step(*pool[i % len(pool)])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        train_step(*pool[i % len(pool)])
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    out = {"ms_per_step": ms_total / K, "value": world * w.B * K / (ms_total * 1e-3), "kernel_ms": kernel_ms,
           "graphs": len(graphs), "collective_in_graph": graph_ok["collective_in_graph"],
           "collective": ("fused into the optimizer step over NVLink peer memory (vmlmf_p2p_adam_step)" if w.bucket.fused_reduce
                          else ("ncclAllReduce of the flat bucket" if world > 1 else None))}

    if want_e2e:
        loader = SyntheticLoader(w.make_batch, dev, pool=4, source="host", seed=1234 + rank)
        if use_graph:                                              # graphs of the two staging buffers, outside the timed region
            for sx, sy in loader.staging_buffers():
                sx.copy_(pool[0][0]); sy.copy_(pool[0][1])
                train_step(sx, sy)

        def e2e_loop(n):
            last = 0.0
            for x, y in loader.batches(n):
                last = train_step(x, y).item()                    # D2H read of the step's result (syncs, as train.py:66 does)
            return last

        e2e_loop(min(W, 5))
        barrier()
        t0 = time.perf_counter()
        e2e_loop(K)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        out["e2e"] = {"value": world * w.B * K / e2e_s, "unit": "sequences/s", "h2d_bytes_per_step": loader.bytes_per_batch,
                      "d2h_bytes_per_step": 4, "ms_per_step": e2e_s * 1e3 / K,
                      "loader": "vmlmf_b200.data.SyntheticLoader(source='host'): pinned pool, double-buffered H2D on a side stream"}

    # inference (no_grad), resident inputs
    w.net.eval()
    for i in range(3):
        w.infer(pool[i % len(pool)][0])
    barrier()
    e0.record()
    for i in range(K):
        w.infer(pool[i % len(pool)][0])
    e1.record()
    barrier()
    inf_ms = max_over_ranks(e0.elapsed_time(e1))
    w.net.train()
    out["inference"] = {"value": world * w.B * K / (inf_ms * 1e-3), "unit": "sequences/s", "ms_per_step": inf_ms / K}
    F.EVENT_LOG = []
    for i in range(3):
        w.infer(pool[i % len(pool)][0])
    torch.cuda.synchronize()
    ilog, F.EVENT_LOG = F.EVENT_LOG, None
    it = {}
    for name, a, b in ilog:
        it.setdefault(name, []).append(a.elapsed_time(b))
    out["infer_kernel_ms"] = {k: sum(v) / 3 for k, v in it.items()}
    return out


def rooflines(w, m):
    """per-call rooflines of a measured workload (train forward / backward calls, inference forward)"""
    a = algorithmic(w.c)
    units = w.B * w.c["T"]
    reg = w.regime()
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        pass
    tr = (traffic or {}) if w.name == "cfg2" else {}
    r_bwd = roofline_of("seq_bwd", m["kernel_ms"].get("seq_bwd"), units, a["q_bwd"], 2 * a["f_step"], reg, tr.get("seq_bwd"))
    r_fwd = roofline_of("seq_fwd", m["kernel_ms"].get("seq_fwd"), units, a["q_fwd"], a["f_step"], reg, tr.get("seq_fwd"))
    r_inf = roofline_of("seq_fwd (inference)", m["infer_kernel_ms"].get("seq_fwd"), units, a["q_inf"], a["f_step"], reg)
    if reg == "R3":
        # small-batch regime: neither HBM nor the tensor pipe binds, the step is a latency chain (two group barriers, two
        # TMA round trips, an L2 reduce); the number that describes it is the time per timestep and layer
        for r in (r_bwd, r_fwd, r_inf):
            if r:
                r["us_per_timestep"] = r["ms_per_launch"] * 1e3 / w.c["T"]
                r["note"] = "latency-bound regime (B <= 32): see us_per_timestep; the roofline fraction is not the limiter"
    return r_bwd, r_fwd, r_inf, reg


def run_ours(args):
    import torch
    import torch.distributed as dist

    from vmlmf_b200 import functional as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything a library prints meanwhile (e.g. NCCL's version banner) goes
    # to stderr; the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the fused path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None       # pinned input pools on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name = args.config
    c = CONFIGS[name]
    B, K, W = (args.batch or c["batch"]), args.steps, max(args.warmup, 3)
    use_graph = not args.no_graph
    sampler = ClockSampler(local) if rank == 0 else None
    w = Workload(name, dev, B, world, rank)
    m = measure(w, K, W, use_graph, True, rank, world, dist)

    # data-parallel sanity: every rank applied the same reduced gradients, so the replicas must still be bit-identical
    in_sync = None
    if world > 1:
        mine = w.opt.pflat.detach().clone()
        ref0 = mine.clone()
        dist.broadcast(ref0, 0)
        diff = (mine != ref0).sum().to(torch.float64)
        dist.all_reduce(diff)
        in_sync = bool(diff.item() == 0)

    # N > 1: the LM data-parallel line BASELINE names (cfg4, weak scaling at 512 sequences per GPU) next to the headline
    lm_dp = None
    if world > 1 and name != "cfg4" and not args.no_configs:
        wl = Workload("cfg4", dev, CONFIGS["cfg4"]["batch"], world, rank)
        ml = measure(wl, 20, 3, use_graph, False, rank, world, dist)
        mine = wl.opt.pflat.detach().clone()
        ref0 = mine.clone()
        dist.broadcast(ref0, 0)
        diff = (mine != ref0).sum().to(torch.float64)
        dist.all_reduce(diff)
        lm_dp = {"workload": CONFIGS["cfg4"]["desc"], "per_gpu_batch": wl.B, "global_batch": wl.B * world, "value": ml["value"],
                 "unit": "sequences/s", "tokens_per_s": ml["value"] * CONFIGS["cfg4"]["T"], "ms_per_step": ml["ms_per_step"],
                 "grad_allreduce_bytes": wl.bucket.nbytes, "replicas_in_sync": bool(diff.item() == 0),
                 "collective_in_graph": ml["collective_in_graph"],
                 "loss_semantics": "local nll_loss = token-mean x local batch; SUM over ranks = token-mean x global batch (lm_test.py:147,153); clip_grad_norm_(5) on the post-allreduce norm (lm_test.py:204)"}
        del wl
        torch.cuda.empty_cache()
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    r_bwd, r_fwd, r_inf, reg = rooflines(w, m)
    cands = [r for r in (r_bwd, r_fwd) if r]
    roofline = dict(max(cands, key=lambda r: r["ms_per_launch"])) if cands else {}
    roofline["note"] = ("one C-ABI call per layer and direction (seq_fwd / seq_bwd: every kernel the call launches); achieved = "
                        "SURVEY 8(d) algorithmic bytes (or flops) / CUDA-event time of the call on eager steps; the timed region "
                        "replays the same step as a CUDA graph.  cfg2's seq_bwd = fused reverse-time recurrence + weight-gradient "
                        "accumulation (tensor-memory accumulators) + partial reduce + the streaming dUx pass.")
    roofline["other_kernels"] = [r for r in (r_bwd, r_fwd, r_inf) if r and r["kernel"] != roofline.get("kernel")]
    roofline["other_kernels"].append({"kernel": "xproj_fwd", "ms_per_launch": m["kernel_ms"].get("xproj_fwd")})
    roofline["step_share"] = {k: v / m["ms_per_step"] for k, v in m["kernel_ms"].items()}
    roofline["regime"] = reg

    # ---------------- per-step recurrence latency at the reference's own batch (cfg1: B=64, T=128) ----------------
    latency = None
    configs = None
    if world == 1:
        import vmlmf_b200 as vb
        torch.manual_seed(3)
        net1 = vb.Net(9, [128], w_rank=8, u_rank=[6], cell=vb.MyVMLMFCell).to(dev)
        x1 = torch.randn(64, 128, 9, device=dev)
        y1 = torch.randint(0, 6, (64,), device=dev)
        for _ in range(5):
            vb.cross_entropy(net1(x1), y1).backward()
        F.EVENT_LOG = []
        for _ in range(20):
            net1.zero_grad()
            vb.cross_entropy(net1(x1), y1).backward()
        torch.cuda.synchronize()
        lat = {}
        for nm, a, b in F.EVENT_LOG:
            lat.setdefault(nm, []).append(a.elapsed_time(b))
        F.EVENT_LOG = None
        latency = {"config": "cfg1 Net(9,[128],8,[6]) B=64 T=128",
                   "fwd_us_per_timestep": statistics.median(lat["seq_fwd"]) * 1e3 / 128,
                   "bwd_us_per_timestep": statistics.median(lat["seq_bwd"]) * 1e3 / 128}

        # ---------------- every BASELINE config, same measurement, fewer steps ----------------
        if not args.no_configs:
            configs = []
            for cn in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg4_b20", "cfg5", "cfg5b"):
                try:
                    if cn == name:
                        wc, mc = w, m
                    else:
                        wc = Workload(cn, dev, CONFIGS[cn]["batch"], 1, 0)
                        big = cn in ("cfg5", "cfg5b", "cfg3")
                        mc = measure(wc, 5 if big else 30, 3, use_graph and cn not in ("cfg5b",), False, 0, 1, dist)
                    rb, rf, ri, rg = rooflines(wc, mc)
                    al = algorithmic(wc.c)
                    configs.append({"name": cn, "workload": wc.c["desc"], "regime": rg, "per_gpu_batch": wc.B, "seq_len": wc.c["T"],
                                    "train": {"value": mc["value"], "unit": "sequences/s", "ms_per_step": mc["ms_per_step"]},
                                    "inference": mc["inference"],
                                    "recurrence_us_per_timestep": {"fwd": (mc["kernel_ms"].get("seq_fwd") or 0) * 1e3 / (wc.c["T"] * wc.c.get("layers", 1)),
                                                                   "bwd": (mc["kernel_ms"].get("seq_bwd") or 0) * 1e3 / (wc.c["T"] * wc.c.get("layers", 1))},
                                    "algorithmic_per_unit": al, "units_per_step": wc.B * wc.c["T"],
                                    "roofline": {"bwd": rb, "fwd": rf, "inference": ri}})
                    if wc is not w:
                        del wc, mc
                        torch.cuda.empty_cache()
                except Exception as e:                             # one config must not take the headline down with it
                    configs.append({"name": cn, "error": f"{type(e).__name__}: {e}"})
                    torch.cuda.empty_cache()

    # ---------------- CPU baseline (bounded sample) ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cb = args.cpu_batch or CPU_BATCH[name]
        rate, ms, done, threads = cpu_train_rate(name, cb, 40, 1, budget_s=20.0)
        cpu = {"value": rate, "unit": "sequences/s", "cores": threads, "kind": "port",
               "sample": f"{done} steps x {cb} sequences of the same {name} train step (oracle port of the reference, torch eager, {ms:.0f} ms/step)"}

    # this library's kernels per step (C-ABI launches; PyTorch adds the ones seed of backward, the multi-tensor gradient
    # gather and the optimizer step counter): R1M cfg2 = pack_plain_fwd, xproj_small, seq_fwd_mma, head_fwd,
    # softmax_nll_fwd + sum_scale, softmax_nll_bwd, head_bwd + head_reduce, seq_bwd_fused, reduce_partials, dux_rows +
    # dux_reduce, pack_plain_bwd, adam = 15.  R2 per layer: pack_plain_fwd, xproj, pack / prep / split / r2_fwd, pack_bwd /
    # r2_bwd, ~14 time-parallel gradient kernels, pack_plain_bwd.
    # R3 per layer: the same minus the x-side split, plus the XP GEMM, vxt_pad, the dzx GEMM.
    per_step = {"R1M": 15, "R1": 13, "R2": 12 + 26 * c.get("layers", 1), "R3": 12 + 28 * c.get("layers", 1),
                "G": 12 + 7 * c["T"] * c.get("layers", 1)}.get(reg, 15)
    out = {
        "metric": METRIC, "value": m["value"], "unit": "sequences/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": c["desc"], "name": name, "per_gpu_batch": B, "global_batch": B * world, "seq_len": c["T"],
                   "parallelism": f"dp{world}", "cuda_graph": bool(use_graph), "cuda_graphs": m["graphs"],
                   "collective_in_graph": m["collective_in_graph"], "collective": m["collective"], "numa_binding": numa,
                   "l2_policy": "inputs larger than L2 (4 rotating resident batches; saved state per step >> 126 MB)"},
        "clocks": clocks,
        "e2e": m.get("e2e"),
        "gpu_launches": per_step * K,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "inference": m["inference"],
        "recurrence_latency": latency,
        "grad_allreduce_bytes": w.bucket.nbytes,
        "replicas_in_sync": in_sync,
        "lm_data_parallel": lm_dp,
        "configs": configs,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
