#!/usr/bin/env python
"""bench.py — scans/s of KissICP::RegisterFrame on KITTI-shape synthetic streams.

Contract (driver):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  * a "step" = one scan (64 x 1024 rays, ~65k points) registered by one pipeline, i.e. one pass
    of the hot path (pipeline/KissICP.cpp:35-68). Workload = BASELINE.json configs[1]
    ("KITTI-00-shape synthetic stream on 1xB200"); at N GPUs every rank runs its own sequence
    (seed = rank) — weak scaling, no collective on the data path, poses gathered at the end.
  * before the W warm-up steps each pipeline is PRIMED with --prime scans (setup, untimed) so the
    timed region sees a steady-state local map instead of an almost empty one.
  * value  = scans/s with the K timed scans already resident in HBM (kb_pipeline_register_frame_dev)
  * e2e    = scans/s through the host-facing C-ABI call (kb_pipeline_register_frame) with pinned
    HOST buffers: H2D of the scan and D2H of the result inside the timed region.
  * --impl reference times the reference's CPU algorithm (the oracle port, OpenMP over the host
    cores; the real TBB/Eigen build is impossible offline — see DESIGN.md) on the same stream.
Timing: CUDA events on the stream the kernels are launched on, barrier + synchronize on both
sides, max over ranks. One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "KITTI-00-shape synthetic stream (64x1024 rays, ~65k pts/scan, voxel 1.0 m, no stamps), 1 sequence per GPU"
METRIC = "scans/sec (65k-pt KITTI-shape clouds)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--prime", type=int, default=100, help="scans registered before warm-up (steady-state map)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=60, help="scans timed by the cpu_baseline leg")
    ap.add_argument("--streams", type=int, default=0, help="extra leg: S independent sequences per GPU on S streams")
    ap.add_argument("--no-nn", action="store_true", help="skip the NN-kernel roofline leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: do not sample clocks (an invalid run by the bench contract)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed regions. In-process NVML (the counters behind
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`): a looping nvidia-smi process needs
    ~0.4 s per sample on an 8-GPU box and was seen stalling 40 ms timed regions; the NVML calls take < 1 ms.
    Falls back to the nvidia-smi loop of B200_PROFILING.md if NVML cannot be used from Python."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index, uuid=None, period=0.05):
        self.index, self.uuid, self.period = index, uuid, period
        self.rows, self.proc, self.nvml, self.first = [], None, None, 0
        self.max_mhz, self.stop_flag, self.source, self.query_ms = None, False, None, []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode())
                except Exception:
                    h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [v for v in vis.split(",") if v.strip()]
                phys = int(ids[self.index]) if ids and all(v.strip().isdigit() for v in ids) else self.index
                h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nvml = (pynvml, h, reasons_fn)
            self.source = "NVML in-process, %d ms period" % int(self.period * 1e3)
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.source = "nvidia-smi -lms 100"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h, reasons_fn = self.nvml
        while not self.stop_flag:
            try:
                t0 = time.perf_counter()
                sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                bits = int(reasons_fn(h))
                self.rows.append((sm, bits))
                self.query_ms.append((time.perf_counter() - t0) * 1e3)
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) > 8 and r[1].replace(".", "").isdigit():
                bits = 0
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        bits |= self.BITS[name]
                if r[2].replace(".", "").isdigit():
                    self.max_mhz = max(self.max_mhz or 0.0, float(r[2]))
                self.rows.append((float(r[1]), bits))

    def wait_ready(self, timeout=20.0):
        """NVML / nvidia-smi start-up stalls the GPU for hundreds of ms: never let it land in a timed region"""
        t0 = time.perf_counter()
        while (self.proc or self.nvml) and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def mark(self):
        self.first = len(self.rows)

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        rows = self.rows[self.first:]
        sm = [r[0] for r in rows]
        reasons = sorted(name for name, bit in self.BITS.items() if any(r[1] & bit for r in rows))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(sm), "source": self.source,
                "query_ms_max": float(max(self.query_ms)) if self.query_ms else None}


# ----------------------------------------------------------------------------- reference arm
def best_thread_count(O, scans, candidates):
    """the oracle's OpenMP regions are tiny (1-2k points per ICP iteration): pick the thread count
    that makes the REFERENCE fastest on this box, so the baseline is not handicapped"""
    best, best_t = None, None
    for nt in candidates:
        icp = O.KissICP(max_num_threads=nt)
        for p, t in scans[:4]:
            icp.register_frame(p, t, want_clouds=False)
        t0 = time.perf_counter()
        for p, t in scans[4:12]:
            icp.register_frame(p, t, want_clouds=False)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = nt, dt
    return best


def thread_candidates():
    n = os.cpu_count() or 1
    c = sorted({min(n, x) for x in (4, 8, 16, 32, 64, n)})
    return c


def run_reference(args, rank, world):
    """the reference's CPU implementation of the path on the host cores (oracle port). At --gpus N the job is
    N independent sequences (one per GPU in our arm): here they run concurrently on the host, each with its
    share of the cores, and value = N * steps / wall."""
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from kiss_icp_b200 import synthetic
    from oracle import oracle as O
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    nseq = max(1, args.gpus)
    n_total = args.prime + args.warmup + args.steps
    streams = []
    for sid in range(nseq):
        lidar = synthetic.kitti_shape(seed=sid, device=dev)
        streams.append([lidar.scan(k) for k in range(n_total)])
    ncpu = os.cpu_count() or 1
    best = best_thread_count(O, streams[0], thread_candidates())
    nt = max(1, min(best, ncpu // nseq))
    icps = [O.KissICP(max_num_threads=nt) for _ in range(nseq)]

    def run(sid, lo, hi):
        for p, t in streams[sid][lo:hi]:
            icps[sid].register_frame(p, t, want_clouds=False)  # ctypes releases the GIL during the call

    with ThreadPoolExecutor(nseq) as ex:
        list(ex.map(lambda sid: run(sid, 0, args.prime + args.warmup), range(nseq)))
        t0 = time.perf_counter()
        list(ex.map(lambda sid: run(sid, args.prime + args.warmup, n_total), range(nseq)))
        dt = time.perf_counter() - t0
    val = nseq * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "scans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "prime_scans": args.prime, "sequences": nseq,
                       "note": "reference CPU path = dependency-free restatement of cpp/kiss_icp (oracle port, OpenMP "
                               "for TBB); the real TBB/Eigen build needs network-fetched deps. N sequences run "
                               "concurrently on the host cores."},
            "cpu_baseline": {"value": val, "unit": "scans/s", "cores": nt * nseq, "kind": "port",
                             "sample": f"{nseq} x {args.steps} scans after {args.prime + args.warmup} untimed; {nt} OpenMP "
                                       f"threads per sequence (fastest single-sequence count of {thread_candidates()} "
                                       f"capped at cpus/sequences) on {ncpu} cpus"},
            "e2e": {"value": val, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def main():
    args = parse()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import kiss_icp_b200 as K
    from kiss_icp_b200 import _native as N, sharding, synthetic

    if not torch.cuda.is_available() or N.lib().kb_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: kiss_icp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N.check(N.lib().kb_set_device(local))
    # torch's default stream has a NULL handle; use an explicit stream so that the library's launches and
    # the torch.cuda.Event timing share ONE stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    N.check(N.lib().kb_set_stream(C.c_void_p(stream.cuda_stream)))

    sampler = ClockSampler(local, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    if rank == 0 and not args.no_clocks:
        sampler.start()  # started long before the timed regions; samples before mark() are dropped

    seq = sharding.sequences_of_rank(rank, world, world)[0]
    lidar = synthetic.kitti_shape(seed=seq, device=dev)
    n_total = args.prime + args.warmup + args.steps
    scans_dev = [lidar.scan_torch(k)[0].contiguous() for k in range(n_total)]
    empty_ts = np.empty(0)
    L = N.lib()

    def make_pipeline():
        return K.KissICP(K.load_config())

    def reg_dev(icp, t):
        N.check(L.kb_pipeline_register_frame_dev(icp._h, C.c_void_p(t.data_ptr()), t.shape[0], None, 0))

    def launches(icp):
        c = C.c_ulonglong(0)
        N.check(L.kb_pipeline_launch_count(icp._h, C.byref(c)))
        return c.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def primed(n_blocking):
        icp = make_pipeline()
        for t in scans_dev[:n_blocking]:
            reg_dev(icp, t)
        return icp

    def queued(icp, tensors, layout):
        """KissICP.register_frames on raw addresses: the whole list is queued on the device, copy of frame k+1
        overlaps the registration of frame k, every frame's result (pose + counters) is read back behind the queue"""
        k = len(tensors)
        return icp._register_frames_raw([t.data_ptr() for t in tensors], [t.shape[0] for t in tensors], [None] * k, [0] * k, layout)

    def timed_region(fn):
        barrier()
        e0.record(stream)
        fn()
        e1.record(stream)
        barrier()
        return sharding.max_over_ranks(e0.elapsed_time(e1), dev)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    warm = scans_dev[args.prime:args.prime + args.warmup]
    timed = scans_dev[args.prime + args.warmup:]
    prof = np.zeros((len(timed), 6))
    iters = np.zeros(len(timed))
    work = np.zeros((len(timed), 2))
    npts = np.zeros(len(timed))
    poses_local = []

    # ---------------- value: inputs resident in HBM, frames queued (kb_pipeline_register_frames, device layout)
    icp = primed(args.prime)
    if rank == 0:
        sampler.wait_ready()  # NVML start-up perturbs the GPU for tens of ms: let it land in the warm-up
    queued(icp, warm, 2)
    if rank == 0:
        sampler.mark()
    l0 = launches(icp)
    ms_max = timed_region(lambda: queued(icp, timed, 2))
    gpu_launches = launches(icp) - l0
    ms_dev = ms_max
    value = world * args.steps / (ms_max * 1e-3)

    # ---------------- the same with one blocking RegisterFrame call per scan (the reference's call shape)
    icp_b = primed(args.prime + args.warmup)
    wall_dev = np.zeros(len(timed))

    def blocking_dev():
        for i, t in enumerate(timed):
            t0 = time.perf_counter()
            reg_dev(icp_b, t)
            wall_dev[i] = time.perf_counter() - t0
    ms_dev_blocking = timed_region(blocking_dev)

    # ---------------- e2e: host-facing call, pinned host buffers, H2D + D2H of every step inside the timed region
    pinned_warm = [t.cpu().pin_memory() for t in warm]
    pinned = [t.cpu().pin_memory() for t in timed]
    h2d = float(np.mean([p.numel() * 8 for p in pinned]))
    icp2 = primed(args.prime)
    queued(icp2, pinned_warm, 0)
    e2e_poses = []
    ms_e2e = timed_region(lambda: e2e_poses.append(queued(icp2, pinned, 0)))
    e2e_value = world * args.steps / (ms_e2e * 1e-3)
    # blocking form: kb_pipeline_register_frame per scan
    icp2b = primed(args.prime + args.warmup)
    wall_e2e = np.zeros(len(pinned))

    def blocking_e2e():
        for i, p in enumerate(pinned):
            t0 = time.perf_counter()
            N.check(L.kb_pipeline_register_frame(icp2b._h, C.c_void_p(p.data_ptr()), p.shape[0], None, 0))
            wall_e2e[i] = time.perf_counter() - t0
    ms_e2e_blocking = timed_region(blocking_e2e)
    # ---------------- profiling pass (untimed): the same scans once more with the in-kernel phase timestamps ON
    # (they cost ~1 us each on the kernel's critical path, so the timed passes above run without them)
    icp4 = primed(args.prime + args.warmup)
    icp4.set_profiling(True)
    icp4.start_history(len(timed))
    for t in timed:
        reg_dev(icp4, t)
    for i, st in enumerate(icp4.history()):
        prof[i] = list(st.phase_us)
        iters[i] = st.iterations
        work[i] = (st.icp_queries, st.icp_candidates)
        npts[i] = st.n_points_in
        poses_local.append(np.array(st.pose).reshape(4, 4))
    # ---------------- e2e, float32 ingestion (KITTI .bin / PointCloud2 are float32; the scans are fp32-representable)
    icp3 = primed(args.prime)
    queued(icp3, [p.to(torch.float32).pin_memory() for p in pinned_warm], 1)
    pinned32 = [p.to(torch.float32).pin_memory() for p in pinned]
    ms_e2e32 = timed_region(lambda: queued(icp3, pinned32, 1))
    same32 = bool(np.array_equal(icp3.last_pose, icp2.last_pose))
    clocks = sampler.stop() if rank == 0 else None
    # ---------------- untimed: the whole trajectory from scan 0 (for the drift metrics below)
    traj = None
    if rank == 0:
        try:
            traj = queued(make_pipeline(), scans_dev, 2)
        except Exception as e:  # never lose the bench line over a side leg
            traj = None
            print("trajectory pass failed:", e, file=sys.stderr)
    d2h = 392.0  # sizeof(FrameResult): pose, delta, sigma, counters, stamps

    # the two pipelines saw identical inputs -> identical trajectories (determinism check)
    same = bool(np.array_equal(icp.last_pose, icp2.last_pose) and np.array_equal(icp_b.last_pose, icp2b.last_pose)
                and np.array_equal(icp.last_pose, icp_b.last_pose) and np.array_equal(e2e_poses[0][-1], icp2.last_pose))
    all_poses = sharding.gather_poses(np.array(poses_local), dev)  # NCCL all_gather of the trajectories

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (k_register_frame, one launch per scan)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    # algorithmic bytes per launch (DESIGN.md "algorithmic bytes"): ICP GetClosestNeighbor traffic
    # (24 + 27*16 + 32 per query + 24 per candidate) + one-off streams (raw scan in, result out)
    bytes_per_launch = work[:, 0] * (24 + 27 * 16 + 32) + 24.0 * work[:, 1] + 24.0 * npts + d2h
    # average launch duration of the dominant kernel: CUDA events over the timed (uninstrumented) region of the
    # resident-input pass, one launch per step, launches queued back to back
    kern_us = np.full(len(timed), ms_dev * 1e3 / len(timed))
    achieved = float(bytes_per_launch.mean() / (kern_us.mean() * 1e-6) / 1e9)
    roofline = {"kernel": "k_register_frame (persistent cooperative, 1 launch/scan)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "algorithmic_bytes_per_launch": float(bytes_per_launch.mean()), "kernel_us": float(kern_us.mean()),
                "traffic": profiled_traffic("r1_register_frame"),
                "traffic_source": "profiles/r1_register_frame_ncu_raw_selected.csv (ncu --set full, bytes per launch)",
                "peak_source": peak_src,
                "note": "latency-bound: ~%.0f ICP iterations x %.0f queries per launch, working set L2-resident; "
                        "kernel time = CUDA-event time of the timed region / launches; see nn_kernel for the bandwidth-bound NN query"
                        % (iters.mean(), work[:, 0].mean() / max(iters.mean(), 1))}

    ms_leg = multi_stream_leg(args, K, N, L, torch, dev, args.streams) if args.streams > 1 else None
    nn = None if args.no_nn else nn_leg(K, N, L, torch, dev, peak)
    cpu = None if args.no_cpu else cpu_leg(args, lidar)
    quality = trajectory_quality(lidar, traj, cpu.pop("poses", None) if cpu else None)

    line = {"metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "prime_scans": args.prime, "sequences": world,
                       "points_per_scan": float(npts.mean()), "icp_iterations_per_scan": float(iters.mean()),
                       "icp_source_points": float((work[:, 0] / np.maximum(iters, 1)).mean()),
                       "icp_candidates_per_query": float(work[:, 1].sum() / max(work[:, 0].sum(), 1.0)),
                       "l2": "every step consumes a new 1.5 MB scan; the local map (the state of the stream) is "
                             "legitimately L2-resident across steps",
                       "phase_us_note": "from a separate untimed pass with in-kernel %globaltimer stamps ON (they add ~1 us per phase/iteration)",
                       "phase_us": dict(zip(["preprocess", "downsample_0.5v", "downsample_1.5v", "icp", "map_update", "epilogue"],
                                            [float(x) for x in prof.mean(0)])),
                       "call_latency_ms": {"resident": {"p50": float(np.percentile(wall_dev, 50) * 1e3), "p99": float(np.percentile(wall_dev, 99) * 1e3), "max": float(wall_dev.max() * 1e3)},
                                           "e2e": {"p50": float(np.percentile(wall_e2e, 50) * 1e3), "p99": float(np.percentile(wall_e2e, 99) * 1e3), "max": float(wall_e2e.max() * 1e3)},
                                           "over_2ms": {"resident": [[int(i), float(wall_dev[i] * 1e3)] for i in np.nonzero(wall_dev > 2e-3)[0][:8]],
                                                        "e2e": [[int(i), float(wall_e2e[i] * 1e3)] for i in np.nonzero(wall_e2e > 2e-3)[0][:8]]}},
                       "deterministic_replay": same, "gathered_trajectories": list(all_poses.shape),
                       "trajectory_quality": quality},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "api": "KissICP.register_frames / kb_pipeline_register_frames: pinned host float64 frames in, poses out; "
                           "frames queued 3 deep (H2D of scan k+1 overlaps registration of scan k)"},
            "blocking_calls": {"note": "one kb_pipeline_register_frame[_dev] call per scan, host waits for every result (the reference's call shape)",
                               "value_resident": world * args.steps / (ms_dev_blocking * 1e-3),
                               "e2e": world * args.steps / (ms_e2e_blocking * 1e-3), "unit": "scans/s"},
            "e2e_f32": {"value": world * args.steps / (ms_e2e32 * 1e-3), "unit": "scans/s", "h2d_bytes_per_step": h2d / 2,
                        "d2h_bytes_per_step": d2h, "same_trajectory_as_f64": same32,
                        "note": "register_frames, float32 host frames (native KITTI/ROS payload), widened on the device"},
            "gpu_launches": int(gpu_launches),
            "roofline": roofline, "nn_kernel": nn, "multi_stream": ms_leg, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def multi_stream_leg(args, K, N, L, torch, dev, S):
    """S independent sequences on ONE GPU: S pipelines, S host threads, S CUDA streams, each persistent grid
    sized to 1/S of the SMs. Reported beside the single-stream headline, never instead."""
    import threading
    from kiss_icp_b200 import synthetic
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    n_total = min(args.prime, 40) + args.steps
    data = []
    for sid in range(S):
        lidar = synthetic.kitti_shape(seed=100 + sid, device=dev)
        data.append([lidar.scan_torch(k)[0].contiguous() for k in range(n_total)])
    torch.cuda.synchronize()
    pipes = [None] * S
    barrier = threading.Barrier(S + 1)
    walls = [0.0] * S

    def worker(sid):
        N.check(L.kb_set_device(dev.index or 0))
        N.check(L.kb_set_stream(None))                 # own non-blocking stream per pipeline
        N.check(L.kb_set_grid_blocks(max(1, sms // S)))
        icp = K.KissICP(K.load_config())
        pipes[sid] = icp
        for t in data[sid][: n_total - args.steps]:
            N.check(L.kb_pipeline_register_frame_dev(icp._h, C.c_void_p(t.data_ptr()), t.shape[0], None, 0))
        barrier.wait()
        t0 = time.perf_counter()
        for t in data[sid][n_total - args.steps:]:
            N.check(L.kb_pipeline_register_frame_dev(icp._h, C.c_void_p(t.data_ptr()), t.shape[0], None, 0))
        walls[sid] = time.perf_counter() - t0
        barrier.wait()

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(S)]
    for t in threads:
        t.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    dt = time.perf_counter() - t0
    for t in threads:
        t.join()
    return {"streams": S, "grid_blocks_per_stream": max(1, sms // S), "scans_per_s": S * args.steps / dt,
            "ms_per_scan_per_stream": float(np.mean(walls)) / args.steps * 1e3,
            "note": "aggregate over S concurrent sequences on one GPU, inputs resident in HBM, host wall clock"}


def profiled_traffic(tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full summary
    (profiles/<tag>_ncu_raw_selected.csv, produced by tools/summarize_profiles.py); None when absent."""
    import csv
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_raw_selected.csv")
    if not os.path.exists(path):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_launch = {}
    for r in csv.DictReader(open(path)):
        if r["metric"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            per_launch[r["launch"]] = per_launch.get(r["launch"], 0.0) + float(r["value"]) * scale.get(r["unit"], 1.0)
    return float(np.mean(list(per_launch.values()))) if per_launch else None


def nn_leg(K, N, L, torch, dev, peak):
    """BASELINE config 5: batched GetClosestNeighbor on a large map — the bandwidth-bound kernel."""
    from kiss_icp_b200 import synthetic
    g = torch.Generator(device="cpu")
    g.manual_seed(5)
    n_raw, n_q = 4_000_000, 1 << 20
    m = K.VoxelHashMap(1.0, 1e9, 20)
    m.add_points(synthetic.surface_cloud(n_raw, seed=5))  # dense surfaces: ~8 points per 1 m voxel like a real map
    stored = torch.from_numpy(m.point_cloud())
    sel = stored[torch.randint(0, stored.shape[0], (n_q,), generator=g)]
    q = (sel + torch.randn(n_q, 3, generator=g, dtype=torch.float64) * 0.3).to(dev).contiguous()
    outp = torch.empty_like(q)
    outd = torch.empty(n_q, dtype=torch.float64, device=dev)
    b = C.c_double(0)
    N.check(L.kb_map_query_bytes_dev(m._h, C.c_void_p(q.data_ptr()), n_q, C.byref(b)))
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2 (126 MB)
    times = []
    for it in range(13):
        flush.fill_(it & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        N.check(L.kb_map_closest_neighbors_dev(m._h, C.c_void_p(q.data_ptr()), n_q, C.c_void_p(outp.data_ptr()),
                                               C.c_void_p(outd.data_ptr())))
        e1.record(stream)
        torch.cuda.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times))
    ach = b.value / (ms * 1e-3) / 1e9
    return {"kernel": "k_nn_query", "map_points": int(stored.shape[0]), "map_voxels": m.num_voxels(), "queries": n_q, "points_per_voxel": float(stored.shape[0]) / max(m.num_voxels(), 1),
            "algorithmic_bytes": b.value, "bytes_per_query": b.value / n_q, "ms": ms, "achieved": ach, "peak": peak,
            "unit": "GB/s", "frac": ach / peak, "traffic": profiled_traffic("r1_nn_query"),
            "l2": "flushed (256 MiB write) before every timed launch", "reps": len(times)}


def trajectory_quality(lidar, traj, cpu_poses):
    """KITTI-style drift of the registered trajectory against the synthetic ground truth (kiss-icp_b200/metrics.py =
    Metrics.cpp restated), for the GPU path and - over the scans the cpu_baseline leg registered - for the CPU port,
    plus the largest pose difference between the two (parity drift is immaterial)."""
    if traj is None:
        return None
    try:
        from kiss_icp_b200 import metrics as M
        n = len(traj)
        gt = np.array([lidar.pose(k) for k in range(n)])
        gt = np.array([np.linalg.inv(gt[0]) @ g for g in gt])  # KISS-ICP poses are relative to the first scan
        dist = float(np.linalg.norm(np.diff(gt[:, :3, 3], axis=0), axis=1).sum())
        out = {"scans": n, "path_m": dist,
               "gpu": dict(zip(["seq_trans_pct", "seq_rot_deg_per_m", "ate_rot_rad", "ate_trans_m"],
                               list(M.sequence_error(gt, traj)) + list(M.absolute_trajectory_error(gt, traj))))}
        if cpu_poses is not None and len(cpu_poses) <= n:
            m = len(cpu_poses)
            out["cpu_port"] = dict(zip(["scans", "seq_trans_pct", "seq_rot_deg_per_m", "ate_rot_rad", "ate_trans_m"],
                                       [m] + list(M.sequence_error(gt[:m], cpu_poses)) + list(M.absolute_trajectory_error(gt[:m], cpu_poses))))
            out["gpu_same_scans"] = dict(zip(["seq_trans_pct", "seq_rot_deg_per_m", "ate_rot_rad", "ate_trans_m"],
                                             list(M.sequence_error(gt[:m], traj[:m])) + list(M.absolute_trajectory_error(gt[:m], traj[:m]))))
            out["max_abs_pose_diff_gpu_vs_cpu_port"] = float(np.abs(traj[:m] - cpu_poses).max())
        return out
    except Exception as e:
        return {"error": repr(e)}


def cpu_leg(args, lidar):
    """oracle (port of the reference CPU path) on the host cores, bounded sample of the same stream."""
    from oracle import oracle as O
    n = min(args.cpu_sample, args.steps)
    scans = [lidar.scan(k) for k in range(args.prime + n)]
    nt = best_thread_count(O, scans, thread_candidates())
    icp = O.KissICP(max_num_threads=nt)
    poses = []
    for p, t in scans[:args.prime]:
        icp.register_frame(p, t, want_clouds=False)
        poses.append(np.array(icp.pose))
    t0 = time.perf_counter()
    for p, t in scans[args.prime:]:
        icp.register_frame(p, t, want_clouds=False)
        poses.append(np.array(icp.pose))
    dt = time.perf_counter() - t0
    return {"poses": np.array(poses), "value": n / dt, "unit": "scans/s", "cores": nt, "kind": "port", "ms_per_scan": dt / n * 1e3,
            "sample": f"{n} scans after {args.prime} untimed priming scans of the same stream (seed 0); OpenMP threads "
                      f"picked as fastest of {thread_candidates()} on {os.cpu_count()} cpus"}


if __name__ == "__main__":
    main()
