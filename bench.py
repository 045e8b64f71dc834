"""Benchmark of the hot path (BASELINE.json metric: frames/s, full detector + decode + RefineNet, 320x240).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the whole path over one batch of synthetic board frames (BASELINE config 3:
batch=256 320x240 per GPU).  N > 1: launched by torchrun, one rank per GPU, frames sharded per rank (weak scaling,
no data-path collective); timing = CUDA events, max over ranks.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec full detector+refine 320x240"
UNIT = "frames/s"
H, W = 240, 320


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs"), bf16_burst=d.get("bf16_tflops"), bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), power_w_max=float(max(power)),
                    samples=len(sm), reasons=sorted(reasons))


def cpu_reference_fps(frames, seconds_budget=20.0, threads=None):
    """The reference's own algorithm on the host cores: the oracle port (torch CPU fp32, same ops and order as
    /root/reference/src/inference.py:32-70), looped like src/benchmark.py:38-53 on a bounded sample."""
    import torch
    import oracle
    from deepcharuco_b200 import weights_io as Wt
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd, sr = Wt.load_state(Wt.DEFAULT_DEEPC), Wt.load_state(Wt.DEFAULT_REFINENET)
    for f in frames[:2]:
        oracle.pipeline.infer_gray(sd, sr, f)          # warm-up
    t0 = time.time()
    done = 0
    while True:
        oracle.pipeline.infer_gray(sd, sr, frames[done % len(frames)])
        done += 1
        if time.time() - t0 > seconds_budget or done >= 4 * len(frames):
            break
    dt = time.time() - t0
    return done / dt, cores, done, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation (oracle port; the Python reference cannot travel to the
    GPU box) on this box's host cores, same metric / config; rank 0 only."""
    if rank != 0:
        return
    from deepcharuco_b200 import synth
    frames = synth.make_frames(32, H, W, seed=1)
    per_step = 8
    import torch
    import oracle
    from deepcharuco_b200 import weights_io as Wt
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd, sr = Wt.load_state(Wt.DEFAULT_DEEPC), Wt.load_state(Wt.DEFAULT_REFINENET)
    i = 0
    for _ in range(args.warmup):
        for _ in range(per_step):
            oracle.pipeline.infer_gray(sd, sr, frames[i % 32]); i += 1
    t0 = time.time()
    for _ in range(args.steps):
        for _ in range(per_step):
            oracle.pipeline.infer_gray(sd, sr, frames[i % 32]); i += 1
    dt = time.time() - t0
    fps = args.steps * per_step / dt
    line = dict(impl="reference", metric=METRIC, value=fps, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=dt / args.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", config=dict(workload=f"batch={args.batch} 320x240 frames, full pipeline (detector + decode + RefineNet)",
                                              note=f"CPU arm: each step is a bounded sample of {per_step} frames of that workload"),
                cpu_baseline=dict(value=fps, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{per_step} frames/step x {args.steps} steps, torch {torch.__version__} CPU, one frame per call as src/benchmark.py"),
                e2e=dict(value=fps, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def other_configs(dc, Nn, synth, deepc, refinenet, eng, peaks, torch):
    """The other BASELINE.json configs on this GPU (device-resident inputs, CUDA events, 2 warm-ups): 1 = one frame per call through
    the drop-in infer_image (src/benchmark.py:38-53), 2 = detector only at batch 64, 4 = RefineNet on 16384 patches, 5 = one GPU's
    shard (256 frames) of the 2048 x 640x480 batch.  frac = algorithmic TFLOP/s / measured sustained bf16 peak."""
    L = Nn.lib()
    peak = peaks["bf16_sustained"]
    out = {}

    def timed(fn, iters):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    # config 1: the reference's own benchmark loop on its sample image (BGR u8, one frame per call, results on the host)
    g = np.load(os.path.join(ROOT, "tests", "golden", "sample_image.npz"))
    img = g["bgr"]
    for _ in range(5):
        kp, _ = dc.infer_image(img, 16, deepc, refinenet, draw_pred=False)
    t0 = time.perf_counter()
    calls = 300
    for _ in range(calls):
        kp, _ = dc.infer_image(img, 16, deepc, refinenet, draw_pred=False)
    dt = time.perf_counter() - t0
    out["1_single_frame_infer_image"] = dict(calls_per_s=calls / dt, ms_per_call=dt / calls * 1e3, corners=int(kp.shape[0]),
                                             matches_reference_golden=bool(np.abs(kp - g["out_refined"]).max() <= 1e-3),
                                             note="src/benchmark.py:38-53 loop through the drop-in infer_image; reference README: > 200 fps on a GTX 1080 Ti")
    # config 2: detector forward only, batch 64
    frames = torch.from_numpy(synth.tile_frames(synth.make_frames(64, seed=1), 64)).cuda()
    loc = torch.empty((64, 65, 30, 40), device="cuda"); ids = torch.empty((64, 17, 30, 40), device="cuda")
    ms = timed(lambda: Nn.check(L.dcu_detector_forward(eng.handle, frames.data_ptr(), 64, loc.data_ptr(), ids.data_ptr(), None)), 10)
    tf = 64 * eng.detector_flops_per_frame() / ms / 1e9
    out["2_detector_b64_320x240"] = dict(ms=ms, frames_per_s=64 / ms * 1e3, tflops_alg=tf, frac=tf / peak)
    # decode + gather as a stand-alone stage on fp32 logits (the HBM-bound kernel of SURVEY 8d)
    counts = torch.empty(64, dtype=torch.int32, device="cuda"); offs = torch.empty(64, dtype=torch.int32, device="cuda")
    tot = torch.zeros(1, dtype=torch.int32, device="cuda"); kpt = torch.empty((eng.max_patches, 4), dtype=torch.int32, device="cuda")
    pt = torch.empty((eng.max_patches, 24, 24), device="cuda")
    ms = timed(lambda: Nn.check(L.dcu_decode_gather(eng.handle, loc.data_ptr(), ids.data_ptr(), frames.data_ptr(), 64, 16, 0, counts.data_ptr(),
                                                    offs.data_ptr(), tot.data_ptr(), kpt.data_ptr(), pt.data_ptr(), None)), 20)
    k2 = int(tot.item())
    nbytes = 64 * 82 * 1200 * 4 + k2 * (2304 + 2304 + 16)
    out["decode_stage_b64"] = dict(ms=ms, gb_per_s_alg=nbytes / ms / 1e6, frac_hbm=nbytes / ms / 1e6 / peaks["hbm"], corners=k2,
                                   note="64 frames = 25 MB of logits: launch-latency dominated at this size")
    # config 4: RefineNet on 16384 real patches (the gather output above, cycled)
    own4 = eng.max_patches < 16384
    e4 = Nn.Engine(deepc._ctx.state_det, deepc._ctx.state_ref, 240, 320, 16, eng.device, max_batch=8, max_patches=16384) if own4 else eng
    reps = (16384 + k2 - 1) // max(k2, 1)
    pt16 = pt[:k2].repeat(reps, 1, 1)[:16384].contiguous(); kp16 = kpt[:k2].repeat(reps, 1)[:16384].contiguous()
    corners = torch.empty((16384, 2), dtype=torch.int32, device="cuda"); refined = torch.empty((16384, 2), device="cuda")
    ms = timed(lambda: Nn.check(L.dcu_refine_forward(e4.handle, pt16.data_ptr(), kp16.data_ptr(), 4, 16384, corners.data_ptr(),
                                                     refined.data_ptr(), None, None)), 5)
    tf = 16384 * e4.refine_flops_per_patch() / ms / 1e9
    out["4_refinenet_16384_patches"] = dict(ms=ms, patches_per_s=16384 / ms * 1e3, tflops_alg=tf, frac=tf / peak)
    if own4:
        e4.close()
    # config 5: one GPU's shard of the 2048 x 640x480 batch (256 frames, 4 boards per frame)
    e5 = Nn.Engine(deepc._ctx.state_det, deepc._ctx.state_ref, 480, 640, 16, eng.device, max_batch=256, max_patches=32768)
    f5 = torch.from_numpy(synth.tile_frames(synth.make_frames(16, 480, 640, seed=1), 256)).cuda()
    ms = timed(lambda: e5.infer_batch_device(f5.data_ptr(), 256, 16, True, None), 3)
    k5 = int(e5._dev_out["total"].item())
    tf = (256 * e5.detector_flops_per_frame() + k5 * e5.refine_flops_per_patch()) / ms / 1e9
    out["5_full_b256_640x480_per_gpu"] = dict(ms=ms, frames_per_s=256 / ms * 1e3, corners=k5, tflops_alg=tf, frac=tf / peak)
    e5.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step (BASELINE config 3: 256)")
    ap.add_argument("--conv", default=os.environ.get("DCU_CONV_IMPL", ""), help="ffma | tcgen05 (default: library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (1, 2, 4, 5) reported next to the headline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        # one slice of the host cores per rank: the ranks' copy / launch threads do not migrate onto each other's cores
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]) or set(cores))
        except (AttributeError, OSError):
            pass
    if args.conv:
        os.environ["DCU_CONV_IMPL"] = args.conv

    import deepcharuco_b200 as dc
    from deepcharuco_b200 import synth, sharding, _native as Nn

    B = args.batch
    deepc, refinenet = dc.load_models(dc.DEFAULT_DEEPC, dc.DEFAULT_REFINENET, n_ids=16, device=local_rank)
    eng = deepc._ctx.engine(H, W, max_batch=B, max_patches=64 * B)
    # synthetic data: a pool of 64 distinct seeded board frames per rank, cycled to the batch; R rotating batches so the
    # input set (R x B x 76.8 kB) exceeds L2; the per-step activation traffic (>100 MB per 4-frame micro-batch) does too.
    pool = synth.make_frames(64, H, W, seed=1 + rank)
    R = 8
    host_batches = []
    for r in range(R):
        fr = synth.tile_frames(np.roll(pool, r * 7, axis=0), B)
        host_batches.append(torch.from_numpy(fr).pin_memory())
    dev_batches = [hb.to(dev, non_blocking=True) for hb in host_batches]
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream

    def step_device(i):
        eng.infer_batch_device(dev_batches[i % R].data_ptr(), B, 16, True, sptr)

    def step_host(i):
        return eng.infer_batch_host(host_batches[i % R].numpy(), 16, True, sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, sampler=None):
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            step_fn(i)
        e1.record(stream)
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), clocks

    for i in range(args.warmup):
        step_device(i)
    l0 = eng.launch_count()
    ms_dev, clocks = timed(step_device, args.steps, ClockSampler(local_rank) if rank == 0 else None)
    launches = eng.launch_count() - l0
    value = world * B * args.steps / (ms_dev / 1e3)

    # roofline of the dominant kernel (3x3 conv): per-launch CUDA events on the launching stream, right after the timed region
    # (same clocks / power state), 2 warm steps in profiling mode + 4 measured steps
    PROF_STEPS = 4
    eng.profile_enable(True)
    for i in range(2):
        step_device(i)
    eng.profile_enable(False)
    eng.profile_enable(True)
    for i in range(PROF_STEPS):
        step_device(i)
    conv_ms, conv_flops, conv_n = eng.profile_read(0)
    conv_issued = eng.profile_read_issued(0)
    dec_ms, dec_bytes, dec_n = eng.profile_read(3)
    first_ms, _, _ = eng.profile_read(1)
    heads_ms, _, _ = eng.profile_read(2)
    eng.profile_enable(False)

    for i in range(max(1, args.warmup // 2)):
        step_host(i)
    ms_e2e, _ = timed(step_host, args.steps)
    counts, offsets, kpts, refined = step_host(0)
    total_k = int(counts.sum())
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = B * H * W
    d2h = 4 + 2 * B * 4 + total_k * (16 + 8)

    # the same through the PYTHON surface a user of the reference switches to: dc.infer_batch (host ndarray in, list of per-frame
    # (K,3) float64 arrays out, incl. the marshalling of inference.py:68-70)
    py_steps = max(3, args.steps // 3)

    def step_python(i):
        return dc.infer_batch(host_batches[i % R].numpy(), 16, deepc, refinenet)
    step_python(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(py_steps):
        step_python(i)
    torch.cuda.synchronize()
    dt_py = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt_py, op=dist.ReduceOp.MAX)
    e2e_python = world * B * py_steps / float(dt_py.item())

    # N > 1: ONE host batch of world x B frames through the user-facing multi-GPU call (every rank holds the batch, runs its
    # shard, results all-gathered over NCCL so every rank returns the full list): scatter + pipeline + merge inside the timed region
    one_batch = None
    if world > 1:
        try:
            big = torch.from_numpy(np.concatenate([hb.numpy() for hb in host_batches[:min(world, R)]] * ((world + R - 1) // R), 0)[:world * B])
            big = big.pin_memory().numpy()              # the caller's batch, page-locked like the per-rank batches above
            dc.infer_batch_distributed(big, 16, deepc, refinenet)
            barrier()
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                full = dc.infer_batch_distributed(big, 16, deepc, refinenet)
            torch.cuda.synchronize()
            dt_ob = torch.tensor([time.perf_counter() - t0], device=dev)
            dist.all_reduce(dt_ob, op=dist.ReduceOp.MAX)
            one_batch = dict(frames=int(world * B), value=world * B * reps / float(dt_ob.item()), unit=UNIT, ms_per_batch=float(dt_ob.item()) / reps * 1e3,
                             api="infer_batch_distributed (same host batch on every rank, own shard per rank, packed results all-gathered over NCCL)",
                             frames_returned=len(full))
            del big
        except Exception as ex:            # never lose the headline line to the optional extra measurement
            one_batch = dict(error=repr(ex)[:300])

    peaks = read_peaks()
    achieved_tf = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    issued_tf = conv_issued / (conv_ms / 1e3) / 1e12 if (conv_ms > 0 and conv_issued > 0) else None
    peak_tf = peaks["bf16_sustained"]
    step_ms_prof = conv_ms + dec_ms + first_ms + heads_ms
    dec_bytes += PROF_STEPS * (total_k * (2304 + 2304 + 16))          # + K*(patch read + patch write + record) per profiled step

    # DRAM traffic of the dominant kernel: from the committed ncu capture of this same workload (per-launch average)
    conv_traffic, traffic_src = None, None
    try:
        ls = json.load(open(os.path.join(ROOT, "profiles", "r2_launch_summary.json")))
        conv_traffic = ls["conv3x3_tc"]["dram_bytes_per_launch"]
        traffic_src = ("profiles/r2_launch_summary.json: ncu dram__bytes_read.sum + dram__bytes_write.sum of THIS workload (tools/profile_step.py "
                       "--batch 256), average per conv_tc2_kernel launch")
    except Exception:
        pass
    if rank == 0:
        impl_name = {Nn.CONV_FFMA: "ffma-fp32", Nn.CONV_TCGEN05: "tcgen05-f16-hi/lo-split(3 products, fp32 accumulate)"}[eng.conv_impl]
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="f32" if eng.conv_impl == Nn.CONV_FFMA else "f16x2-split(f32-equivalent)", data="synthetic",
            config=dict(workload=f"batch={B} 320x240 frames per GPU, full pipeline (detector + decode + RefineNet)",
                        frames_per_gpu_per_step=B, corners_per_step_rank0=total_k, conv_impl=impl_name,
                        weights="trained reference checkpoints (converted .npz)",
                        l2="inputs rotate over 8 distinct batches (157 MB > 126 MB L2); per-step activation traffic >> L2",
                        parallelism=f"replicas x{world}, batch-sharded, no data-path collective"),
            clocks=clocks,
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                     ms_per_step=ms_e2e / args.steps, api="dcu_infer_batch_host (pinned host u8 frames in, packed keypoints out)"),
            e2e_python=dict(value=e2e_python, unit=UNIT, steps=py_steps,
                            api="deepcharuco_b200.infer_batch (host ndarray in, list of per-frame (K,3) float64 arrays out)"),
            gpu_launches=int(launches),
            roofline=dict(bound="tensor", kernel="conv_tc2_kernel (CTA-pair tcgen05 3x3 convolution; all 3x3 layer shapes of both networks, aggregated over launches)",
                          achieved=achieved_tf, peak=peak_tf, unit="TFLOP/s", frac=achieved_tf / peak_tf if peak_tf else None,
                          peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
                          note="achieved = ALGORITHMIC flops (2*MAC of the reference's fp32 convolutions, SURVEY.md 8d).  The parity-safe "
                               "fp16 hi/lo split issues 3 tensor-core products per MAC (ceiling: frac 1/3); upsample-fused layers issue 4 of 9 "
                               "taps.  issued_tflops = what the tensor pipes execute (all MMAs, padded tile rows included).",
                          issued_tflops=issued_tf,
                          issued_frac=(issued_tf / peak_tf) if (peak_tf and issued_tf) else None,
                          traffic=conv_traffic, traffic_source=traffic_src, launches=int(conv_n // PROF_STEPS), kernel_ms_per_step=conv_ms / PROF_STEPS,
                          share_of_step=conv_ms / step_ms_prof if step_ms_prof else None,
                          decode_gather=dict(bound="hbm", logit_read="eliminated", launches=int(dec_n // PROF_STEPS), kernel_ms_per_step=dec_ms / PROF_STEPS,
                                             actual_bytes_per_step=int(B * 1200 * 2 + total_k * (2304 + 2304 + 16)),
                                             survey_bytes_per_step=int(dec_bytes / PROF_STEPS),
                                             note="the per-cell arg-max is taken in the 1x1 head epilogues, so the 82 logit planes of SURVEY 8d "
                                                  "(393.6 kB / frame) are neither written nor re-read: the kernel that is left reads 2 B per cell + the "
                                                  "patch windows and is launch-latency bound (no GB/s figure is meaningful); the stand-alone stage "
                                                  "dcu_decode_gather on fp32 logits is timed under configs.decode_stage")),
        )
        if one_batch is not None:
            line["one_batch"] = one_batch
        if not args.no_configs and world == 1:
            line["configs"] = other_configs(dc, Nn, synth, deepc, refinenet, eng, peaks, torch)
        if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only (the other ranks would be spinning on the barrier)
            fps, cores, done, dt = cpu_reference_fps(pool[:32], seconds_budget=15.0)
            line["cpu_baseline"] = dict(value=fps, unit=UNIT, cores=cores, kind="port",
                                        sample=f"{done} frames of the same synthetic set in {dt:.1f} s, one frame per call "
                                               f"(src/benchmark.py loop), torch CPU fp32 oracle")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
