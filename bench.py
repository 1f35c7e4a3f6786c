#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native DeepSets hot path.

Metric (BASELINE.json): train samples/sec (fwd+bwd), DeepSets B=256 N=1024 per GPU.
A "step" = forward + BCEWithLogitsLoss + backward of one batch (+ gradient all-reduce
when N > 1) of the workload BASELINE.json's configs[1] names: phi [3->256->256]+final,
ReLU, max pool, rho [256]->10, bf16 tensor-core path, 1xB200 (weak scaling for N > 1:
256 sets per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W]           # this repo (CUDA)
  python bench.py --impl reference [...]                         # reference CPU path (oracle port)
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank per GPU

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "point-cloud-classifier_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_PER_GPU, N_PTS, D_IN, H, OUT = 256, 1024, 3, 256, 10
CFG = dict(input_dim=D_IN, phi_layers=[H, H], rho_layers=[H], output_dim=OUT, activation="relu", layer_norm=False,
           residual_block=False, pooling="max")
# algorithmic FLOP per point, reference formulation, recompute not credited
FLOP_FWD_PT = 2 * (H * D_IN + 2 * H * H)               # 263,680
FLOP_TRAIN_PT = 3 * FLOP_FWD_PT - 2 * H * D_IN          # 789,504
FLOP_CHAIN_PT = 2 * (2 * H * H)                         # dgrad of the two H x H layers
FLOP_WGRAD_PT = FLOP_FWD_PT                             # wgrad of all three layers
N_ROTATE = 32                                           # distinct input batches (166 MB > 126 MB L2)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                "hbm": float(d.get("hbm_gbs", 6650.0)), "source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML from a background thread, ~2 ms
    period; the timed region lasts tens of milliseconds, too short for `nvidia-smi -lms`)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop, self._thr, self._err = False, None, None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            while not self._stop:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for k, b in bits.items():
                        if r & b:
                            self.reasons.add(k)
                except Exception:
                    pass
                time.sleep(0.002)
        except Exception as e:  # NVML missing: report it, never fail the bench
            self._err = f"{type(e).__name__}: {e}"

    def start(self):
        import threading
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop = True
        if self._thr is not None:
            self._thr.join(timeout=2)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self._err:
            out["error"] = self._err
        return out


def make_batches(n_batches, B, device, seed):
    g = torch.Generator().manual_seed(seed)
    host = []
    idx = torch.arange(B).repeat_interleave(N_PTS)
    for _ in range(n_batches):
        x = torch.randn(B * N_PTS, D_IN, generator=g)
        y = (torch.rand(B, OUT, generator=g) > 0.5).float()
        host.append((x.pin_memory() if device != "cpu" else x, idx.pin_memory() if device != "cpu" else idx,
                     y.pin_memory() if device != "cpu" else y))
    return host


# ---------------------------------------------------------------------------- reference arm
def cpu_reference_rate(steps, warmup, sample_sets=32, budget_s=25.0):
    """The reference's own CPU path (oracle/deepsets_oracle.py: functional restatement of
    models/deep_sets.py + wrapper.py:38 loss, pinned to the reference by tests/golden) on the
    box's host cores.  One step = fwd + loss + bwd of a `sample_sets`-set sample of the workload."""
    from oracle import deepsets_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.init_state_dict(CFG, seed=0)
    (x, idx, y), = make_batches(1, sample_sets, "cpu", seed=1)
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.deepsets_train_step(sd, CFG, x, idx, y)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin > budget_s and len(times) >= 2:
            break
    ms = statistics.median(times) * 1e3
    return {"value": sample_sets / ms * 1e3, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} steps of fwd+loss+bwd on {sample_sets} sets x {N_PTS} pts (same model), median",
            "ms_per_step": ms, "steps": len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_rate(args.steps, args.warmup, budget_s=120.0)
    line = {"impl": "reference", "metric": "train samples/sec (fwd+bwd), DeepSets B=256 N=1024", "value": cb["value"],
            "unit": "samples/s", "n_gpus": args.gpus, "steps": cb["steps"], "warmup": args.warmup,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "DeepSets phi[3-256-256]+final relu max-pool rho[256]-10, fwd+loss+bwd",
                       "device": "host CPU", "sample_sets_per_step": 32, "points_per_set": N_PTS},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------- this repo
def run_ours(args):
    import torch.distributed as dist
    import pcc_b200
    from pcc_b200 import _lib
    from pcc_b200.train_step import GraphedTrainStep
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a); there is no CPU fallback. Use --impl reference for the CPU arm.")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    def log(msg):
        if rank == 0:
            print(f"[bench] {msg}", file=sys.stderr, flush=True)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        log(f"process group up: world {world}")
    _lib.call("pcc_check_device", local)

    torch.manual_seed(0)
    model = pcc_b200.DeepSets(**CFG, precision=args.precision).to(dev)
    host = make_batches(N_ROTATE, B_PER_GPU, "cuda", seed=1000 + rank)
    devb = [tuple(t.to(dev) for t in b) for b in host]
    kw = {"num_sets": B_PER_GPU}

    def build(use_graph):
        return GraphedTrainStep(model, devb[0][:2], devb[0][2], forward_kwargs=kw, allreduce=world > 1,
                                use_graph=use_graph)
    try:
        gs = build(not args.no_graph)
        graphed = not args.no_graph
    except Exception as e:  # graph capture unavailable (e.g. NCCL capture) -> eager launches
        if rank == 0:
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running eager", file=sys.stderr)
        gs = build(False)
        graphed = False
    assert model.last_path == ("fused-bf16" if args.precision == "bf16" else "fp32")
    log(f"train step built (cuda graph: {graphed})")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    W = max(args.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # Both timed loops feed the captured step the same way, like a prefetching input pipeline: two captured steps
    # (GraphedTrainStep instances over the same model) own one set of static input buffers each; while step i
    # computes out of slot i%2, the copy stream moves batch i+1 straight into the other slot's static buffers.
    #   value: the 32 rotating batches are already resident in HBM (device -> device copies);
    #   e2e:   they sit in pinned host memory (H2D inside the timed region) and the loss of every step is copied
    #          back to pinned memory on a third stream and read by the host one step later (wrapper.py:73).
    copy_stream = torch.cuda.Stream()
    d2h_stream = torch.cuda.Stream()
    slots = [gs, build(graphed)]

    def make_feed(batches, readback):
        staged_ev = [torch.cuda.Event() for _ in range(2)]
        free_ev = [torch.cuda.Event() for _ in range(2)]
        loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]
        state = {"primed": False, "seen": 0}

        def enqueue_copy(i):
            b, slot = batches[i % N_ROTATE], i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free_ev[slot])          # the step that consumed this slot has finished with it
                slots[slot].load(b[:2], b[2])                  # batch -> the slot's static device buffers
                staged_ev[slot].record(copy_stream)

        def step(i):
            main = torch.cuda.current_stream()
            if not state["primed"]:
                for ev in free_ev:
                    ev.record(main)
                enqueue_copy(i)
                state["primed"] = True
            slot = i % 2
            main.wait_event(staged_ev[slot])
            loss = slots[slot].run()
            free_ev[slot].record(main)
            enqueue_copy(i + 1)                                # overlaps with this step's compute
            if readback:
                with torch.cuda.stream(d2h_stream):            # off the compute stream: the next replay does not
                    d2h_stream.wait_event(free_ev[slot])       # queue behind a copy-engine round trip
                    loss_host[slot].copy_(loss.detach(), non_blocking=True)
                    loss_ev[slot].record(d2h_stream)
                if state["seen"] > 0:
                    loss_ev[1 - slot].synchronize()            # host reads the previous step's loss
                    _ = float(loss_host[1 - slot])
            state["seen"] += 1
        return step

    # ---- value: inputs resident in HBM when the timed region starts
    ms_dev = timed(make_feed(devb, False), args.steps, W)
    torch.cuda.synchronize()
    # ---- e2e: pinned host buffers -> H2D inside the timed region, loss read back every step
    e2e_step = make_feed(host, True)
    ms_e2e = timed(e2e_step, args.steps, W)
    log(f"e2e timing done: {ms_e2e:.3f} ms/step")
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel durations (CUDA events inside the library, eager launches of the same step)
    eager = GraphedTrainStep(model, devb[0][:2], devb[0][2], forward_kwargs=kw, allreduce=False, use_graph=False, warmup=2)
    _lib.call("pcc_launch_count", 1)
    eager.step(devb[1][:2], devb[1][2])
    torch.cuda.synchronize()
    launches = int(_lib.call("pcc_launch_count", 1))
    _lib.call("pcc_prof_enable", 1)
    ksteps = min(args.steps, 20)
    for i in range(ksteps):
        eager.step(devb[i % N_ROTATE][:2], devb[i % N_ROTATE][2])
    torch.cuda.synchronize()
    _lib.call("pcc_prof_enable", 0)
    kern = {}
    for slot, name in ((0, "phi_pool_fwd_kernel"), (1, "phi_bwd_chain_kernel"), (2, "phi_wgrad_kernel")):
        ms_tot, cnt = C.c_double(0), C.c_int64(0)
        _lib.call("pcc_prof_read", slot, C.byref(ms_tot), C.byref(cnt))
        if cnt.value:
            kern[name] = ms_tot.value / cnt.value
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    pts = B_PER_GPU * N_PTS
    flops = {"phi_pool_fwd_kernel": FLOP_FWD_PT, "phi_bwd_chain_kernel": FLOP_CHAIN_PT, "phi_wgrad_kernel": FLOP_WGRAD_PT}
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):   # dram bytes per launch from the committed `ncu --set full` capture (tools/summarize_profile.py)
        traffic = json.load(open(tpath)).get("kernels", {})
    roof = None
    if kern:
        dom = max(kern, key=kern.get)
        ach = flops[dom] * pts / (kern[dom] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["tflops"], "traffic": traffic.get(dom), "traffic_unit": "dram bytes per launch (ncu --set full, profiles/ncu_traffic.json)",
                "peak_source": peaks["source"],
                "kernel_ms": {k: round(v, 4) for k, v in kern.items()},
                "algorithmic_flop_per_point": flops[dom]}
    step_tf = FLOP_TRAIN_PT * pts / (ms_dev * 1e-3) / 1e12
    cb = cpu_reference_rate(8, 2) if world == 1 else None
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    line = {
        "metric": "train samples/sec (fwd+bwd), DeepSets B=256 N=1024", "value": world * B_PER_GPU / ms_dev * 1e3,
        "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms_dev,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: DeepSets B=256 N=1024 per GPU, phi[3-256-256]+final(256) relu, max pool, "
                               "rho[256]-10, fwd + BCEWithLogitsLoss + bwd" +
                               ((" + gradient all-reduce (own one-shot kernel over NVLink peer memory, inside the graph)"
                                 if gs.peer is not None else " + NCCL gradient all-reduce") if world > 1 else ""),
                   "sets_per_gpu": B_PER_GPU, "points_per_set": N_PTS, "cuda_graph": graphed,
                   "l2": f"inputs rotate over {N_ROTATE} distinct batches (166 MB) and every step streams ~0.7 GB of "
                         "staged operands, both larger than the 126 MB L2",
                   "parallelism": f"dp{world}", "precision_mode": args.precision},
        "e2e": {"value": world * B_PER_GPU / ms_e2e * 1e3, "unit": "samples/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "api": "GraphedTrainStep.load/run, two captured slots fed alternately from pinned host x, idx, y (H2D on a copy stream, one step ahead) + per-step loss read-back on a third stream; the device-resident value uses the same feed with device-to-device copies"},
        "gpu_launches": launches * args.steps,
        "gpu_launches_per_step": launches,
        "roofline": roof,
        "roofline_step": {"bound": "tensor", "achieved": step_tf, "peak": peaks["tflops"], "unit": "TFLOP/s",
                          "frac": step_tf / peaks["tflops"], "algorithmic_flop_per_point": FLOP_TRAIN_PT},
        "cpu_baseline": ({k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")} if cb else None),
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
