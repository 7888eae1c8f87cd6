#!/usr/bin/env python
"""bench.py — throughput of the photometric-reconstruction hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3]

One "step" = one pass of the hot path over one synthetic batch: Trainer.pred_novel_images +
Trainer.compute_losses (photometric term) forward AND backward (trainer.py:523-603, 701-742), at
BASELINE.json configs[1]: batch 12/GPU, 640x192, 49 vertical planes, stereo disparity warp,
0.85*SSIM + 0.15*L1.  The perceptual network and the smoothness term are out of the path's scope
(cuDNN / a 1-plane stencil) and are not executed by either arm.

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM (CUDA-graph replay of
the step); `e2e` = same metric through the public API with the `inputs` dict in pinned HOST memory
(H2D per step, as trainer.py:328-329 does) and a D2H read of the loss per step; `roofline` = achieved
algorithmic HBM GB/s of the dominant kernel (CUDA events on the launch stream) against
MEASURED_PEAKS.json; `cpu_baseline` = the same path on the host cores (bounded sample).
`--impl reference` times the reference's own CPU implementation (baseline/_ref through an import shim
when that directory travelled with the snapshot, else the oracle port).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # name: (B per GPU, H, W, opt overrides, photometric mode, description)
    "cfg2": (12, 192, 640, dict(), "ssim_l1", "cfg2: B=12/GPU 640x192 N=49 stereo disp_warp + SSIM/L1, fwd+bwd"),
    "cfg3": (4, 384, 1280, dict(use_mixture_loss=True, plane_residual=True), None,
             "cfg3: B=4/GPU 1280x384 N=49 disp_warp + plane_residual + Laplacian mixture, fwd+bwd"),
    # per-GPU shapes of the multi-GPU BASELINE configs (parity / diagnostics; the bench line is cfg2)
    "cfg4": (12, 192, 640, dict(warp_type="homography_warp", xz_levels=14, novel_frame_ids=[-1, 1], automask=True), None,
             "cfg4 (per GPU): B=12 640x192 N=49+14 homography_warp, frames [r,-1,1], automask L1, fwd+bwd"),
    "cfg5": (8, 384, 1280, dict(use_mixture_loss=True, plane_residual=True, xz_levels=14, self_distillation=1.0), None,
             "cfg5 (per GPU): B=8 1280x384 N=49+14 disp_warp + mixture + mask_novel blend, fwd+bwd"),
    # diagnostics (not BASELINE configs)
    "dbg_mix": (4, 384, 1280, dict(use_mixture_loss=True), None, "debug: mixture only"),
    "dbg_res": (4, 384, 1280, dict(plane_residual=True), None, "debug: residual only"),
    "dbg_w": (4, 384, 1280, dict(), None, "debug: 1280 wide L1"),
}
RAW_H, RAW_W = 375, 1242  # a KITTI raw frame (the reference's dataset), what a decoder hands to the loader
METRIC = "training images/sec (49-plane warp+SSIM hot path, fwd+bwd)"
UNIT = "images/s"


# --------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions: an NVML polling thread (the timed regions last
    milliseconds, far below nvidia-smi's start-up time) plus inline samples taken by the main thread while it waits for
    the queued steps to drain.  Falls back to `nvidia-smi -lms` when NVML cannot be loaded."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.sm, self.bits, self.smax, self.h, self.nv = [], 0, None, None, None
        self.running = False

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    idx = self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.smax = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.h = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def sample(self):
        """One NVML sample (callable from the main thread while the GPU works)."""
        if self.h is None:
            return
        try:
            self.sm.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            try:
                self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass

    def sample_until(self, event):
        """Sample while `event` (recorded after the queued steps) has not completed."""
        while not event.query():
            self.sample()

    def _poll(self):
        while self.running:
            self.sample()
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.h is not None:
            self.running = False
            self.thread.join(timeout=1)
            sm = sorted(self.sm)
            reasons = [n for n, bit in self.REASONS.items() if self.bits & bit]
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": reasons, "samples": len(sm),
                    "source": "nvml, polled during the timed regions"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi -lms 20"}


def gpu_cpu_affinity(index):
    """CPUs NVML reports as local to the GPU (its NUMA node), or None.  Pinned staging buffers are first-touched by the thread
    that allocates them: allocating from a CPU of the GPU's node keeps the H2D DMA off the inter-socket link."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = index
        if vis:
            try:
                idx = int(vis.split(",")[index])
            except Exception:
                idx = index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


def algorithmic_bytes(B, N, H, W, mixture, elem=4):
    """SURVEY.md §8(d) per image per target frame; X1 = H*W*4; `elem` = bytes per stored logit / sigma / gradient element
    (4, or 2 with bf16 storage, which halves the N-terms).
    warp fwd: N(1+m) logits/sigma + 3 src + 3 rgb_rec (+3 tgt, +1 nll with mixture)
    warp bwd: 2N(1+m) (re-read, write grads) + 3 src + 3 g_rgb_rec + 3 rgb_rec/tgt
    loss fwd: 3 rgb_rec + 3 tgt (+1 ph map) ; loss bwd: 3 rgb_rec + 3 tgt + 3 g_rgb_rec."""
    x1 = H * W * 4
    m = 1 if mixture else 0
    return {
        "pd_warp_composite_fwd": B * x1 * (N * (1 + m) * elem / 4 + 6 + 4 * m),
        "pd_warp_composite_bwd": B * x1 * (2 * N * (1 + m) * elem / 4 + 9),
        "pd_photometric_fwd": B * x1 * 6,
        "pd_photometric_bwd": B * x1 * 9,
    }


# --------------------------------------------------------------------------------------------------
# reference / CPU arm
# --------------------------------------------------------------------------------------------------
def load_reference(cpu_only=True):
    """The unmodified reference through the import shim (baseline/shim.py), if baseline/_ref travelled with the snapshot.
    Returns (trainer module, layers module) or None."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import shim

        mods = shim.load(cpu_only=cpu_only)
        return None if mods is None else (mods[0], mods[1])
    except Exception as e:  # pragma: no cover
        sys.stderr.write("reference import failed (%s); using the oracle port\n" % e)
        return None


def cpu_step_fn(cfg_name, B_sample, seed=0):
    """Returns (callable doing one fwd+bwd of the path on CPU for B_sample images, kind)."""
    from types import SimpleNamespace

    import torch.nn as nn

    from planedepth_b200.synthetic import make_batch, make_opt

    _, H, W, over, photometric, _ = CONFIGS[cfg_name]
    opt = make_opt(**over)
    batch = make_batch(B_sample, H, W, opt, seed=seed, device="cpu")
    leaves = list(batch.leaves.values())
    ref = load_reference()
    if ref is not None:
        ref_trainer, ref_layers = ref
        t = object.__new__(ref_trainer.Trainer)
        t.opt = SimpleNamespace(**vars(opt))
        t.opt.use_ssim = photometric == "ssim_l1"
        t.target_sides = batch.target_sides
        t.softmax = nn.Softmax(1)
        t.ssim = ref_layers.SSIM()
        t.homography_warp = ref_layers.HomographyWarp(H, W)
        t.backproject_depth = ref_layers.BackprojectDepth(H, W)
        t.project_3d = ref_layers.Project3D(H, W)

        def step():
            out = batch.attach(dict(batch.outputs))
            ref_trainer.Trainer.pred_novel_images(t, batch.inputs, out)
            total = 0
            for s in batch.target_sides:
                if photometric == "ssim_l1":
                    total = total + ref_trainer.Trainer.compute_reprojection_loss(t, out[("rgb_rec", s)], batch.inputs[("color", s)]).mean()
                elif opt.use_mixture_loss:
                    err = torch.abs(out[("rgb_rec_layered", s)] - batch.inputs[("color", s)][:, None]).mean(2)
                    total = total + ref_layers.multimodal_loss(err, out[("sigma_rec", s)], out[("pi_rec", s)], dist="lap").mean()
                else:
                    total = total + torch.abs(out[("rgb_rec", s)] - batch.inputs[("color", s)]).mean()
            torch.autograd.grad(total, leaves, allow_unused=True)
            return float(total)

        return step, "reference"
    from oracle import pd_oracle as O

    def step():
        out = batch.attach(dict(batch.outputs))
        O.pred_novel_images(opt, batch.target_sides, batch.inputs, out)
        total = 0
        for s in batch.target_sides:
            ph, _ = O.photometric_map(opt, batch.inputs, out, s, photometric)
            total = total + ph.mean()
        torch.autograd.grad(total, leaves, allow_unused=True)
        return float(total)

    return step, "port"


def time_cpu(cfg_name, B_sample, steps, warmup):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = cpu_step_fn(cfg_name, B_sample)
    try:
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
    finally:
        # the CPU arm runs the reference with `.cuda()` patched to the identity (baseline/shim.py): undo it for the GPU legs
        try:
            import shim

            shim.leave_cpu_mode()
        except Exception:
            pass
    return {"value": B_sample / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d step(s) of B=%d images of %s (fwd+bwd), %d torch threads" % (steps, B_sample, cfg_name, cores)}, dt


def run_reference(args):
    from planedepth_b200 import dist as D

    rank, _, ws = D.env()
    if rank != 0:
        return
    B, H, W, over, photometric, desc = CONFIGS[args.config]
    # each step is a bounded sample of the workload (images of the configured shape), sized so that K steps of the
    # CPU path (about 0.1 s per image on 16 host threads at cfg2) end within a few minutes
    B_sample = min(B, 4 if args.steps <= 30 else (2 if args.steps <= 100 else 1))
    cb, dt = time_cpu(args.config, B_sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "sample_batch": B_sample, "device": "host CPU"},
        "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from planedepth_b200 import _lib, functional
    from planedepth_b200 import dist as D
    from planedepth_b200.boundary import HotPath
    from planedepth_b200.graph import GraphedStep, make_step
    from planedepth_b200.synthetic import make_batch, make_opt

    rank, local_rank, ws = D.env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if ws > 1:
        # NCCL_DEBUG is left as the launcher set it (its INFO lines carry the rank / transport evidence); protect_stdout()
        # keeps fd 1 to the one JSON line whatever NCCL prints
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.lib()
    if args.no_fuse_bwd:
        functional.FUSE_PHOTOMETRIC_BWD = False
    dev = torch.device("cuda", local_rank)

    def measure(cfg_name, steps, warmup, main):
        B, H, W, over, photometric, desc = CONFIGS[cfg_name]
        opt = make_opt(**over)
        mnov = opt.self_distillation > 0  # the SD stage hands over outputs["mask_novel"] (trainer.py:404-466)
        opt.self_distillation = 0.0      # its |disp - disp_pp| term belongs to the decoder tail, not to this path
        seed = D.shard_seed(1234, rank)
        batch = make_batch(B, H, W, opt, seed=seed, device="cpu", layout=args.layout, mask_novel=mnov)
        N = batch.shape[1]
        dev = torch.device("cuda", local_rank)
        # static device buffers: network outputs live on the device; the `inputs` dict comes from the host.  The pinned staging
        # buffers are allocated (first-touched) from a CPU of the GPU's NUMA node
        old_aff, local_cpus = os.sched_getaffinity(0), (None if args.no_numa_pin else gpu_cpu_affinity(local_rank))
        if local_cpus:
            os.sched_setaffinity(0, local_cpus)
        host_inputs = {k: v.pin_memory() for k, v in batch.inputs.items()}
        batch_gpu = make_batch(B, H, W, opt, seed=seed, device=dev, layout=args.layout, mask_novel=mnov)
        outputs, leaves_map = batch_gpu.outputs, batch_gpu.leaves
        inputs = batch_gpu.inputs
        if args.storage == "bf16":
            # secondary configuration: the network outputs are stored as bf16 (pd_warp_desc.dtype), gradients come back as bf16
            for k in ("logits", "sigma"):
                if k in leaves_map:
                    leaves_map[k] = leaves_map[k].detach().to(torch.bfloat16).requires_grad_(True)
                    outputs[k] = leaves_map[k]
            outputs["probability"] = outputs["logits"].detach()
        leaves = list(leaves_map.values())
        # the integration patch promises x-constant disparities whenever the decoder has no yz planes
        hp = HotPath(opt, batch.target_sides, pc_net=None, photometric=photometric, disp_rowwise=(getattr(opt, "yz_levels", 0) == 0) and not args.no_rowwise)
        step = make_step(hp, inputs, outputs, leaves, batch_gpu.attach)
        if args.no_graph:
            class _Eager:
                result = None

                def replay(self):
                    self.result = step()
                    return self.result

            graphed = _Eager()
            graphed.replay()
        else:
            graphed = GraphedStep(step, warmup=3)

        def barrier():
            D.barrier(ws, cuda=True)

        # ---------------- value: inputs resident in HBM, CUDA-graph replay -------------------------
        for _ in range(warmup):
            graphed.replay()
        clocks = ClockSampler(local_rank)
        barrier()
        lib.pd_reset_launch_count()
        probe = step()  # one eager step to count the library launches a step performs
        launches_per_step = lib.pd_launch_count()
        del probe
        barrier()
        if rank == 0:
            clocks.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            graphed.replay()
        ev1.record()
        if rank == 0:
            clocks.sample_until(ev1)  # the queued steps are still running: these samples are under load
        barrier()
        ms_step = D.max_over_ranks(ev0.elapsed_time(ev1), ws, dev) / steps
        loss_val = float(graphed.result["loss"].item())

        # ---------------- e2e: the path's host-born inputs from pinned host memory each step ---------
        # The tensors of the `inputs` dict this path reads (source + target colours; K / inv_K / Rt for the
        # homography warp) start every step in pinned HOST memory, as they do when they come out of the data loader
        # (trainer.py:328-329 copies them with .to(device)); logits / sigma / plane geometry are network outputs and are
        # device-born in the reference too.  The H2D copy of step i+1 runs on a copy stream while step i computes
        # (one staging set, events both ways); the loss comes back to the host every step.
        color = "color"
        read_keys = [(color, "l")] + [(color, sd) for sd in batch.target_sides]
        if opt.warp_type != "disp_warp":
            read_keys += ["K", "inv_K"] + [("Rt", sd) for sd in batch.target_sides]
        read_keys = [k for k in dict.fromkeys(read_keys) if k in host_inputs]
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        copy_stream = torch.cuda.Stream()

        def run_e2e(u8):
            """u8: the colour frames start as RAW uint8 frames of the dataset's size (KITTI: 375 x 1242 x 3, interleaved, as a
            decoder leaves them) and are converted / resized / clamped on the device (pd_resize_bicubic_u8)
            — what the reference's loader does per sample on the CPU (pair_transforms.py:63-78) before shipping fp32 tensors.
            Otherwise: the fp32 tensors of the `inputs` dict travel, as in the reference (trainer.py:328-329)."""
            host_sel = {}
            for k in read_keys:
                if u8 and isinstance(k, tuple) and k[0] == color:
                    gg = torch.Generator().manual_seed(seed + 17 + len(host_sel))
                    host_sel[k] = torch.randint(0, 256, (B, RAW_H, RAW_W, 3), generator=gg, dtype=torch.uint8).pin_memory()
                else:
                    host_sel[k] = host_inputs[k]
            # two staging sets: the H2D of step i+1 only waits for the consumer of step i-1 (uint8 frames are converted /
            # resized by a kernel on the main stream; with one set the next transfer would idle for that kernel's duration.
            # Converting on the copy stream instead was measured and dropped: its CTAs only find SMs between the persistent
            # kernels of the step, 1.7 ms per step instead of 0.7)
            staging = [{k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host_sel.items()} for _ in range(2)]
            nbytes = sum(v.numel() * v.element_size() for v in host_sel.values())
            copied = [torch.cuda.Event(), torch.cuda.Event()]
            consumed = [torch.cuda.Event(), torch.cuda.Event()]
            state = {"next_h2d": 0, "next_use": 0}

            def enqueue_h2d():
                i = state["next_h2d"]
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[i])
                    for k, v in host_sel.items():
                        staging[i][k].copy_(v, non_blocking=True)
                    copied[i].record(copy_stream)
                state["next_h2d"] = i ^ 1

            for ev in consumed:
                ev.record()
            enqueue_h2d()
            enqueue_h2d()

            def e2e_step():
                main = torch.cuda.current_stream()
                i = state["next_use"]
                main.wait_event(copied[i])
                for k in read_keys:
                    if staging[i][k].dtype == torch.uint8:
                        functional.resize_frames_u8(staging[i][k], (H, W), out=inputs[k])
                    else:
                        inputs[k].copy_(staging[i][k], non_blocking=True)
                consumed[i].record(main)
                state["next_use"] = i ^ 1
                enqueue_h2d()  # refill the set just consumed: travels while this step (and the next) compute
                res = graphed.replay()
                loss_host.copy_(res["loss"], non_blocking=True)
                main.synchronize()
                return float(loss_host)

            for _ in range(warmup):
                e2e_step()
            barrier()
            ev0.record()
            t0 = time.perf_counter()
            for _ in range(steps):
                e2e_step()
            ev1.record()
            barrier()
            ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
            ms = D.max_over_ranks(ms, ws, dev) / steps
            copy_stream.synchronize()
            return ms, nbytes

        # both transports are measured in every run; the headline e2e is the faster one unless a flag pins it.  (Equal bytes at
        # 640x192: the fp32 tensors are the steadier of the two there; 4.2x fewer bytes for uint8 at 1280x384.)
        ms_f32, h2d_f32 = run_e2e(u8=False)
        ms_u8, h2d_u8 = run_e2e(u8=True)
        e2e_u8 = (not args.e2e_fp32) and (args.e2e_u8 or ms_u8 <= ms_f32)
        (e2e_ms_step, h2d), (e2e_alt_ms, h2d_alt) = ((ms_u8, h2d_u8), (ms_f32, h2d_f32)) if e2e_u8 else ((ms_f32, h2d_f32), (ms_u8, h2d_u8))
        # restore the resident inputs for the roofline pass below
        for k in read_keys:
            inputs[k].copy_(host_inputs[k].to(dev))
        clk = clocks.stop() if rank == 0 else None

        # ---------------- roofline: per-kernel CUDA events on the launch stream (eager pass) --------
        functional.KERNEL_TIMELINE = []
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        functional.KERNEL_TIMELINE = []
        n_roof = max(3, min(steps, 10))
        for _ in range(n_roof):
            # The launches of an eager step are issued by Python far slower than the GPU executes them: an event pair
            # around a call would also time the idle GPU waiting for the host to reach the launch.  A short device-side
            # delay lets the host queue the whole step first, so every event pair brackets nothing but its kernels.
            torch.cuda._sleep(4_000_000)  # ~2 ms at 1.97 GHz
            step()
            torch.cuda.synchronize()
        per = {}
        for name, s, e in functional.KERNEL_TIMELINE:
            per.setdefault(name, []).append(s.elapsed_time(e))
        functional.KERNEL_TIMELINE = None
        avg_ms = {k: sum(v) / len(v) for k, v in per.items()}
        dom = max(avg_ms, key=avg_ms.get)
        alg = algorithmic_bytes(B, N, H, W, opt.use_mixture_loss, 2 if args.storage == "bf16" else 4)  # per launch (= per target side)
        n_calls = {k: len(v) / n_roof for k, v in per.items()}  # launches per step
        peak, peak_src = peaks()
        achieved = alg[dom] / (avg_ms[dom] * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                key = cfg_name + ("_compact" if args.layout == "compact" else ("_nopromise" if args.no_rowwise else ""))
                traffic = tj.get(key, {}).get(dom)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "frac_of_spec_8TBs": achieved / 8000.0, "algorithmic_bytes": alg[dom], "kernel_ms": avg_ms[dom],
                    "all_kernels_ms": avg_ms,
                    "all_kernels_frac": {k: alg[k] / (avg_ms[k] * 1e-3) / 1e9 / peak for k in avg_ms},
                    "launches_per_step": n_calls,
                    "step_frac": sum(alg[k] * n_calls.get(k, 0) for k in alg) / (ms_step * 1e-3) / 1e9 / peak}

        if local_cpus:
            os.sched_setaffinity(0, old_aff)
        fmt = {True: "raw uint8 frames %dx%dx3 converted / bicubic-resized / clamped on the device (pd_resize_bicubic_u8)" % (RAW_H, RAW_W),
               False: "fp32 tensors of the inputs dict (the reference's transport, trainer.py:328-329)"}
        e2e_scope = ("per step: H2D of the path's host-born inputs (%s) from pinned memory on a copy stream overlapped with the previous "
                     "step [colour frames: %s], D2H of the loss; logits/sigma/plane geometry are network outputs (device-born)" % (
                         ", ".join(str(k) for k in read_keys), fmt[e2e_u8]))
        if rank != 0:
            return None
        cpu = None  # filled in after the GPU legs (run_ours)
        return {
            "metric": METRIC, "value": D.aggregate_throughput(B, ws, ms_step), "unit": UNIT, "n_gpus": ws, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "storage": "bf16 logits / sigma / gradients, fp32 arithmetic" if args.storage == "bf16" else "fp32", "batch_per_gpu": B, "global_batch": B * ws, "planes": N, "height": H, "width": W,
                       "photometric": photometric or ("mixture" if opt.use_mixture_loss else "l1"), "layout": args.layout,
                       "rowwise_promise": bool((getattr(opt, "yz_levels", 0) == 0) and not args.no_rowwise),
                       "parallelism": "dp%d (independent shards; the path itself has no collective - see the ddp leg)" % ws,
                       "launch": "eager" if args.no_graph else "cuda_graph replay",
                       "l2": "no flush: per-step working set (logits %.0f MB + grads %.0f MB) exceeds the 126 MB L2" % (
                           B * N * H * W * 4 / 1e6, B * N * H * W * 4 / 1e6),
                       "e2e_scope": e2e_scope},
            "e2e": {"value": D.aggregate_throughput(B, ws, e2e_ms_step), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms_step, "h2d_GBps": h2d / (e2e_ms_step * 1e-3) / 1e9,
                    "numa_pinned_cpus": (len(local_cpus) if local_cpus else 0), "colour_transport": "uint8 raw frames" if e2e_u8 else "fp32",
                    "transport_choice": ("--e2e-fp32" if args.e2e_fp32 else ("--e2e-u8" if args.e2e_u8 else "the faster of the two transports measured in this run")),
                    "other_transport": {"colour_transport": "fp32" if e2e_u8 else "uint8 raw frames",
                                        "value": D.aggregate_throughput(B, ws, e2e_alt_ms), "ms_per_step": e2e_alt_ms, "h2d_bytes_per_step": h2d_alt}},
            "gpu_launches": int(launches_per_step) * steps,
            "gpu_launches_per_step": int(launches_per_step),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clk, "loss": loss_val,
        }

    line = measure(args.config, args.steps, args.warmup, main=True)
    extras = {}
    # the multi-GPU BASELINE configs on the GPU counts they are quoted on (configs[3]: 4 GPUs, configs[4]: 8 GPUs)
    if ws == 4 and args.config == "cfg2":
        extras["cfg4_dp4"] = brief(measure("cfg4", 30, 5, main=False))
    if ws == 8 and args.config == "cfg2":
        extras["cfg5_dp8"] = brief(measure("cfg5", 20, 5, main=False))
    if not args.no_ddp_leg:
        extras["ddp"] = ddp_leg(args, rank, local_rank, ws, dev)
    if ws == 1 and not args.no_reference_gpu:
        extras["reference_gpu"] = reference_gpu_leg(args.config)
    if rank == 0 and ws == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload on the host cores: ~10-20 s of CPU work (26 steps of B=4 images at ~0.35 s).  Last:
        # the CPU arm patches `.cuda()` to the identity while it runs the reference
        line["cpu_baseline"], _ = time_cpu(args.config, min(CONFIGS[args.config][0], 4), 24, 2)
    if rank == 0:
        line.update(extras)
        emit(line)
    if ws > 1:
        dist.destroy_process_group()


def ddp_leg(args, rank, local_rank, ws, dev):
    """The data-parallel TRAINING step north_star describes: a producer network (the reference's own seeded ResNet-18
    encoder + DepthDecoder from baseline/_ref; a built-in convolutional head of the same gradient size when that tree did not
    travel) feeds the hot path, its backward feeds DistributedDataParallel, whose bucketed NCCL all-reduce (sum / world
    over NVLink) overlaps the remaining backward — the only cross-device traffic (no SyncBatchNorm).  cfg2 shape, eager
    launches, device-timed, max over ranks.  At N = 1 the same step runs without the wrapper: the per-N values of this leg
    are what a DDP scaling efficiency is computed from."""
    import torch.distributed as dist
    import torch.nn as nn

    from planedepth_b200 import dist as D
    from planedepth_b200.boundary import HotPath
    from planedepth_b200.synthetic import make_batch, make_opt

    B, H, W, over, photometric, desc = CONFIGS["cfg2"]
    opt = make_opt(**over)
    batch = make_batch(B, H, W, opt, seed=D.shard_seed(4321, rank), device=dev, requires_grad=False)
    inputs = batch.inputs
    N = batch.shape[1]
    producer = None
    kind = "builtin conv head"
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import shim

        if shim.available() and shim.load() is not None:
            ropt = shim.default_options(num_layers=18, height=H, width=W, xz_levels=0, novel_frame_ids=[])
            models = shim.build_models(ropt, dev, seed=7)

            class RefProducer(nn.Module):
                def __init__(self):
                    super().__init__()
                    self.encoder, self.depth = models["encoder"], models["depth"]

                def forward(self, img, grid):
                    return self.depth(self.encoder(img), grid)

            producer = RefProducer().to(dev)
            kind = "reference ResnetEncoder(18) + DepthDecoder (baseline/_ref, seeded random weights)"
    except Exception as e:  # pragma: no cover
        sys.stderr.write("ddp leg: reference networks unavailable (%s); built-in head\n" % e)
        producer = None
    if producer is None:
        torch.manual_seed(7)

        class Head(nn.Module):
            def __init__(self):
                super().__init__()
                ch = [3, 64, 128, 256, 512]
                down = []
                for i in range(4):
                    down += [nn.Conv2d(ch[i], ch[i + 1], 3, 2, 1), nn.ELU(inplace=True)]
                mid = []
                for _ in range(5):
                    mid += [nn.Conv2d(512, 512, 3, 1, 1), nn.ELU(inplace=True)]
                self.down, self.mid = nn.Sequential(*down), nn.Sequential(*mid)
                self.out = nn.Conv2d(512, N, 3, 1, 1)

            def forward(self, img, grid):
                x = self.out(self.mid(self.down(img)))
                logits = nn.functional.interpolate(x, size=img.shape[-2:], mode="bilinear", align_corners=False)
                return {"logits": logits, "probability": logits.detach(), "disp": logits[:, :1].abs() + 1.0}

        producer = Head().to(dev)
    hp = HotPath(opt, batch.target_sides, pc_net=None, photometric=photometric, disp_rowwise=True)

    def run_step(net):
        out = net(inputs[("color_aug", "l")], inputs["grid"])
        for k in ("disp_layered", "padding_mask", "distance", "norm"):
            out.setdefault(k, batch.outputs[k])
        out[("Rt", "r")] = inputs[("Rt", "r")]
        return hp.process(inputs, out)["loss/total_loss"]

    # parameters the step never touches (the ResNet's classifier head: the reference copes with
    # find_unused_parameters=True, trainer.py:99, a per-step graph traversal) are frozen once, found by a dry run: DDP then
    # runs with its static bucket plan
    D.freeze_unused_parameters(producer, lambda: run_step(producer))
    n_params = sum(p.numel() for p in producer.parameters() if p.requires_grad)
    model = producer
    if ws > 1:
        model = nn.parallel.DistributedDataParallel(producer, device_ids=[local_rank], output_device=local_rank, gradient_as_bucket_view=True)
    params = [p for p in producer.parameters() if p.requires_grad]

    def step():
        loss = run_step(model)
        for p in params:
            p.grad = None
        loss.backward()
        return loss.detach()

    steps = max(5, min(args.steps, 30))
    for _ in range(5):
        step()
    D.barrier(ws, cuda=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    D.barrier(ws, cuda=True)
    ms = D.max_over_ranks(e0.elapsed_time(e1), ws, dev) / steps
    return {"value": D.aggregate_throughput(B, ws, ms), "unit": UNIT, "ms_per_step": ms, "steps": steps, "n_gpus": ws, "producer": kind,
            "parameters": n_params, "gradient_bytes_allreduced_per_step": (n_params * 4 if ws > 1 else 0),
            "collective": ("DistributedDataParallel bucketed NCCL all-reduce, overlapped with backward" if ws > 1 else "none (N = 1)"),
            "workload": "producer fwd + " + desc + " + producer bwd", "loss": float(loss)}


def reference_gpu_leg(cfg_name):
    """BASELINE.md row G, the denominator of north_star's '>= 10x the reference GPU grid_sample + SSIM path': the UNMODIFIED
    reference code (baseline/_ref; F.grid_sample + autograd) doing the work of one bench step — pred_novel_images + the
    photometric term, forward + backward — on the same synthetic batch on this GPU.  3 warm-up + 10 timed steps."""
    import torch.nn as nn

    from planedepth_b200.synthetic import make_batch, make_opt

    ref = load_reference(cpu_only=False)
    if ref is None:
        return {"unavailable": "baseline/_ref did not travel with the snapshot"}
    tr, layers = ref
    from types import SimpleNamespace

    B, H, W, over, photometric, desc = CONFIGS[cfg_name]
    opt = make_opt(**over)
    mnov = opt.self_distillation > 0
    opt.self_distillation = 0.0
    try:
        batch = make_batch(B, H, W, opt, seed=1234, device="cuda", layout="reference", mask_novel=mnov)
        leaves = list(batch.leaves.values())
        t = object.__new__(tr.Trainer)
        t.opt = SimpleNamespace(**vars(opt))
        t.opt.use_ssim = photometric == "ssim_l1"
        t.target_sides = batch.target_sides
        t.softmax = nn.Softmax(1)
        t.ssim = layers.SSIM().cuda()
        t.homography_warp = layers.HomographyWarp(H, W)
        t.backproject_depth = layers.BackprojectDepth(H, W)
        t.project_3d = layers.Project3D(H, W)

        def step():
            out = batch.attach(dict(batch.outputs))
            tr.Trainer.pred_novel_images(t, batch.inputs, out)
            total = 0
            for s in batch.target_sides:
                if photometric == "ssim_l1":
                    total = total + tr.Trainer.compute_reprojection_loss(t, out[("rgb_rec", s)], batch.inputs[("color", s)]).mean()
                elif opt.use_mixture_loss:
                    err = torch.abs(out[("rgb_rec_layered", s)] - batch.inputs[("color", s)][:, None]).mean(2)
                    total = total + layers.multimodal_loss(err, out[("sigma_rec", s)], out[("pi_rec", s)], dist="lap").mean()
                else:
                    total = total + torch.abs(out[("rgb_rec", s)] - batch.inputs[("color", s)]).mean()
            torch.autograd.grad(total, leaves, allow_unused=True)
            return total

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            loss = step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        res = {"value": B / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "steps": 10, "warmup": 3, "loss": float(loss.detach()),
               "what": "unmodified reference Trainer.pred_novel_images + photometric term (F.grid_sample, autograd), fwd+bwd, same batch, this GPU",
               "warped_images_per_s": B * batch.shape[1] * len(batch.target_sides) / ms * 1e3}
        del batch, leaves, t
        torch.cuda.empty_cache()
        return res
    except Exception as e:  # pragma: no cover
        return {"unavailable": "reference GPU run failed: %s" % (str(e)[:200],)}


def brief(line):
    if line is None:
        return None
    keep = ("value", "unit", "n_gpus", "ms_per_step", "steps", "e2e", "loss", "gpu_launches_per_step")
    out = {k: line[k] for k in keep}
    out["workload"] = line["config"]["workload"]
    out["roofline_frac"] = line["roofline"]["frac"]
    out["roofline_kernel"] = line["roofline"]["kernel"]
    return out


_JSON_FD = None


def protect_stdout():
    """stdout carries exactly ONE JSON line.  Libraries write there behind Python's back (NCCL prints its version banner on
    fd 1 even at NCCL_DEBUG=WARN), so fd 1 is pointed at stderr for the whole run and the line goes to a saved duplicate."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--layout", default="reference", choices=["reference", "compact"])
    ap.add_argument("--no-rowwise", action="store_true",
                    help="withhold the integrator's promise that plane geometry is x-constant (yz_levels == 0): dense masks are streamed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-fuse-bwd", action="store_true", help="keep pd_photometric_bwd as its own launch (diagnostic)")
    ap.add_argument("--storage", default="fp32", choices=["fp32", "bf16"],
                    help="bf16: logits / sigma and their gradients are stored as bf16 (secondary configuration; arithmetic stays fp32)")
    ap.add_argument("--e2e-fp32", action="store_true", help="headline e2e ships fp32 colour tensors (the reference's transport); default: the faster of the two transports")
    ap.add_argument("--e2e-u8", action="store_true", help="headline e2e ships raw uint8 frames resized on the device; default: the faster of the two transports")
    ap.add_argument("--no-numa-pin", action="store_true", help="do not move the process to the GPU's NUMA node for the e2e leg")
    ap.add_argument("--no-ddp-leg", action="store_true", help="skip the producer + DistributedDataParallel training-step leg")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip timing the unmodified reference on this GPU (N = 1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
