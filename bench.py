#!/usr/bin/env python
"""bench.py -- Pointnet2Backbone forward throughput (scenes/s) on synthetic ScanNet-shaped scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): B=8 scenes per GPU, 40 000 points, xyz + height + 128-d
multiview features, SA1-SA4 + FP1-FP2, fused MLPs.  A step = one backbone forward over one batch.
Scenes are sharded by rank with no collective on the data path (weak scaling).

One JSON line on rank 0:
  value     whole-job scenes/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the public API with HOST (pinned) inputs: H2D of the batch and D2H
            of fp2_features/fp2_xyz/fp2_inds inside the timed region, double-buffered
  roofline  the dominant kernel of the step (largest share of device time), timed live with CUDA
            events; roofline_kernels lists every kernel of the step the same way
  cpu_baseline  the CPU oracle (oracle/) on this box's host cores, bounded sample, rank 0 only
  reference_cuda  the reference's own CUDA kernels (oracle/_ref) + stock PyTorch modules on the
            same GPU, when that extension is present -- the existing GPU implementation

--impl reference times the reference's path on the host CPU (the reference ships no CPU kernels,
its ops assert "CPU not supported", so this is the oracle port: same wiring in PyTorch CPU ops over
the C restatement of its kernels, all host threads).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Pointnet2Backbone scenes/sec (40k pts)"
UNIT = "scenes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=None, help="fp32 | bf16 (default: bf16 when built, else fp32)")
    ap.add_argument("--batch", type=int, default=8, help="scenes per GPU per step")
    ap.add_argument("--points", type=int, default=40000)
    ap.add_argument("--lanes", type=int, default=8,
                    help="batches in flight per GPU: step i runs on CUDA stream i %% lanes (1 = strictly serial steps)")
    ap.add_argument("--fps-sms", type=int, default=0,
                    help="give the sampling chains their own group of >= this many SMs (CUDA green contexts); 0 = off")
    ap.add_argument("--no-graphs", action="store_true",
                    help="enqueue every step from Python instead of replaying a captured CUDA graph per lane")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-breakdown", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end loop")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm": float(p["hbm_gbs"]), "tensor_burst": float(p["bf16_tflops"]),
                "tensor": float(p["bf16_tflops_sustained"]), "source": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock and throttle reasons of the samples taken inside [t_begin, t_end] (wall clock);
        all samples when the window caught fewer than two."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                why = [nm for nm, v in zip(names, r[4:8]) if v.strip().lower().startswith("active")]
                parsed.append((ts, float(r[1]), float(r[2]), why))
            except Exception:
                continue
        inside = [q for q in parsed if t_begin is not None and t_begin <= q[0] <= t_end]
        use = inside if len(inside) >= 2 else parsed
        if use:
            out["sm_mhz"] = statistics.median([q[1] for q in use])
            out["sm_max_mhz"] = use[-1][2]
            out["samples"] = len(use)
            out["window"] = "timed regions" if use is inside else "whole run (timed regions shorter than the sampling period)"
            out["reasons"] = sorted({w for q in use for w in q[3]})
        return out


# ---------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's path on the host CPU (oracle port), bounded sample per step."""
    if rank != 0:
        return
    import torch
    from oracle import pn2_oracle as orc
    from situation3d_b200.backbone_module import Pointnet2Backbone
    from situation3d_b200.synthetic import make_batch, randomize_bn_stats
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = randomize_bn_stats(Pointnet2Backbone(input_feature_dim=129)).eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    scenes = 2                                      # bounded sample: 2 of the step's 8 scenes
    pc = torch.from_numpy(make_batch(scenes, args.points, 129))
    with torch.no_grad():
        t0 = time.perf_counter()
        orc.backbone(pc, sd)                        # first call: page-in + thread pools
        first = time.perf_counter() - t0
        steps = max(1, min(args.steps, int(150.0 / max(first, 1e-3))))
        warm = min(args.warmup, 2)
        for _ in range(warm):
            orc.backbone(pc, sd)
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.backbone(pc, sd)
        dt = time.perf_counter() - t0
    value = scenes * steps / dt
    sample = "%d scenes of %d points per step x %d steps (of the B=%d step); oracle port: PyTorch CPU wiring over " \
             "oracle/pn2_oracle.c (OpenMP), %d torch threads" % (scenes, args.points, steps, args.batch, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Pointnet2Backbone forward, %d-point ScanNet-shaped scenes, 129 feature "
                                   "channels, SA 2048/1024/512/256 + 2 FP (BASELINE configs[1] shape)" % args.points,
                       "scenes_per_step": scenes, "device": "host CPU"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def kernel_breakdown(net, pc, precision, pk, reps=5):
    """Every kernel of one step launched alone through the C ABI and timed with CUDA events on the
    launching (current) stream; algorithmic bytes / FLOPs per launch as defined in DESIGN.md."""
    import torch
    from situation3d_b200 import fused
    B, N, W = pc.shape
    C = W - 3
    sas = (net.sa1, net.sa2, net.sa3, net.sa4)
    imgs = net._fused_images(pc)
    xyz = pc[..., :3].contiguous()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    flush = torch.empty(1024 * 1024 * 1024 // 4, dtype=torch.float32, device=pc.device)

    def timeit(fn):
        fn()
        ts = []
        for _ in range(reps):
            # 1 GB write: evicts L2 between repetitions and keeps the GPU busy (~170 us) while the host
            # enqueues the launch, so the events bracket device time, not launch latency
            flush.zero_()
            s, e = ev(), ev()
            s.record()
            fn()
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
        return statistics.median(ts) * 1e-3

    rows = []
    src_xyz, table, ld, c, skip = xyz, pc[..., 3:], W, C, 3
    feats_rows = []
    cxyz_all = []
    for lvl, m in enumerate(sas):
        n_in = src_xyz.shape[1]
        name = "sa%d" % (lvl + 1)
        inds, cxyz = fused.fps_with_xyz(src_xyz, m.npoint)
        t = timeit(lambda: fused.fps_with_xyz(src_xyz, m.npoint))
        # FPS is a chain of dependent rounds: besides the (meaningless) HBM figure, report rounds/s and how much of the
        # machine's FP32 issue rate the distance updates use (4 instructions per point-update: 3 packed sub/mul/fma
        # pairs + min; 148 SMs x 128 lanes)
        upd = B * (m.npoint - 1) * n_in / t
        rows.append({"kernel": "fps_" + name, "bound": "hbm", "seconds": t,
                     "alg_bytes": B * (12 * n_in + 16 * m.npoint), "rounds_per_s": (m.npoint - 1) / t,
                     "point_updates_per_s": upd, "fp32_issue_frac": upd * 4 / (148 * 128 * 1.965e9)})
        idx = fused.ball_query(src_xyz, cxyz, m.radius, m.nsample)
        t = timeit(lambda: fused.ball_query(src_xyz, cxyz, m.radius, m.nsample))
        rows.append({"kernel": "ball_query_" + name, "bound": "hbm", "seconds": t,
                     "alg_bytes": B * (12 * (n_in + m.npoint) + 4 * m.npoint * m.nsample)})
        inv_r = 1.0 / m.radius
        dims = imgs[lvl].dims
        pairs = B * m.npoint * m.nsample
        flops = 2 * pairs * sum(a * b for a, b in zip(dims[:-1], dims[1:]))          # the reference's 1x1 convs
        wbytes = 4 * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
        abytes = B * (4 * (c + 3) * min(n_in, m.npoint * m.nsample) + 4 * m.npoint * m.nsample
                      + 2 * 4 * dims[-1] * m.npoint) + wbytes
        if precision == "bf16" and not imgs[lvl].f32_only:
            # the tensor-core arm is two launches: the row table (bf16 pack, or the per-point half of layer 1)
            # and the fused gather + MLP + max-pool kernel
            prep = lambda: fused.sa_bf16_table(imgs[lvl], table, ld, c, B, n_in, m.npoint, m.nsample, skip)
            tab, image, c_eff = prep()
            split = c_eff != c
            t_prep = timeit(prep) if (split or table.dtype != torch.bfloat16) else 0.0
            out = torch.empty((B, dims[3], m.npoint), dtype=torch.float32, device=pc.device)
            out_rows = torch.empty((B, m.npoint, dims[3]), dtype=torch.bfloat16, device=pc.device)
            run = lambda: fused.sa_bf16_fused(dims, image, c_eff, src_xyz, cxyz, idx, tab, inv_r, True, out, out_rows)
            run()
            t = timeit(run)
            lin_flops = 2 * pairs * c * dims[1] if split else 0      # reference FLOPs the per-point GEMM stands for
            if t_prep > 0:
                in_bytes = B * n_in * (ld * 4 if table.dtype != torch.bfloat16 else tab.shape[2] * 2)
                rows.append({"kernel": "%s_%s" % (name, "pointwise_l1" if split else "pack_rows"), "bound": "hbm",
                             "seconds": t_prep, "alg_bytes": in_bytes + tab.numel() * 2})
            rows.append({"kernel": "%s_fused_bf16" % name, "bound": "tensor", "seconds": t,
                         "alg_flops": flops - lin_flops, "alg_bytes": abytes,
                         "layer_tflops": flops / (t + t_prep) / 1e12})
        else:
            run = lambda: fused.SA_FORWARD[precision](imgs[lvl], src_xyz, cxyz, idx, table, ld, c, True, inv_r, raw_skip=skip)
            out, out_rows = run()
            t = timeit(run)
            rows.append({"kernel": "%s_fused_%s" % (name, precision), "bound": "tensor", "seconds": t,
                         "alg_flops": flops, "alg_bytes": abytes})
        feats_rows.append(out_rows)
        cxyz_all.append(cxyz)
        src_xyz, table, ld, c, skip = cxyz, out_rows, out_rows.shape[2], out_rows.shape[2], 0
    known_rows = feats_rows[3]
    for name, (u, k), skip_rows in (("fp1", (2, 3), feats_rows[2]), ("fp2", (1, 2), feats_rows[1])):
        un, kn = cxyz_all[u], cxyz_all[k]
        d2, i3 = fused.three_nn(un, kn)
        t = timeit(lambda: fused.three_nn(un, kn))
        rows.append({"kernel": "three_nn_" + name, "bound": "hbm", "seconds": t,
                     "alg_bytes": B * (12 * (un.shape[1] + kn.shape[1]) + 24 * un.shape[1])})
        img = imgs[4 if name == "fp1" else 5]
        run = lambda: fused.FP_FORWARD[precision](img, d2, i3, known_rows, skip_rows)
        _, out_rows = run()
        t = timeit(run)
        dims = img.dims
        n = un.shape[1]
        flops = 2 * B * n * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
        abytes = B * (4 * known_rows.shape[2] * kn.shape[1] + 4 * skip_rows.shape[2] * n + 24 * n + 2 * 4 * dims[-1] * n) \
            + 4 * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
        rows.append({"kernel": "%s_fused_%s" % (name, precision), "bound": "tensor", "seconds": t,
                     "alg_flops": flops, "alg_bytes": abytes})
        known_rows = out_rows
    total = sum(r["seconds"] for r in rows)
    for r in rows:
        r["share"] = r["seconds"] / total
        if r["bound"] == "tensor":
            r["achieved"], r["unit"], r["peak"] = r["alg_flops"] / r["seconds"] / 1e12, "TFLOP/s", pk["tensor_burst"]
            r["hbm_gbs"] = r["alg_bytes"] / r["seconds"] / 1e9
        else:
            r["achieved"], r["unit"], r["peak"] = r["alg_bytes"] / r["seconds"] / 1e9, "GB/s", pk["hbm"]
        r["frac"] = r["achieved"] / r["peak"]
        r["us"] = r.pop("seconds") * 1e6
    return rows


def ncu_traffic():
    """DRAM bytes per launch of the bench kernels from the committed ncu --set full captures
    (profiles/ncu_traffic.json: kernel -> dram__bytes_read.sum + dram__bytes_write.sum)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return {}


def reference_cuda_arm(pc, sd, steps=3):
    """The reference's own CUDA kernels (oracle/_ref, unmodified sources) wired as its Python modules
    wire them (oracle restatement) with stock PyTorch conv/BN/ReLU/max-pool on the same GPU."""
    import torch
    from oracle import pn2_oracle as orc
    from oracle.build_ref import load_ref_ext
    ext = load_ref_ext()
    if ext is None:
        return None
    sd = {k: v.to(pc.device) for k, v in sd.items()}
    with torch.no_grad():
        orc.backbone(pc, sd, ops=ext)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            orc.backbone(pc, sd, ops=ext)
        e.record()
        e.synchronize()
    dt = s.elapsed_time(e) * 1e-3
    return {"value": pc.shape[0] * steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "what": "reference _ext_src kernels compiled unmodified for sm_100 + stock torch 1x1 conv/BN/ReLU/"
                    "max_pool2d (cuDNN), fp32 (TF32 conv default), B=%d, 1 GPU" % pc.shape[0]}


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer is allocated, so
    that the staging buffers of the end-to-end loop are first-touched on the GPU's NUMA node (with 8 ranks the H2D
    copies otherwise cross the socket interconnect).  Returns the previous affinity (restored for the CPU baseline)."""
    try:
        import pynvml
        import torch
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local_rank)
        try:
            bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1} & prev
        if cpus:
            os.sched_setaffinity(0, cpus)
        return prev, len(cpus)
    except Exception:
        return None, 0


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from situation3d_b200 import fused
    from situation3d_b200.backbone_module import Pointnet2Backbone
    from situation3d_b200.synthetic import make_batch, randomize_bn_stats

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    prev_affinity, numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else (None, 0)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    precision = args.precision or ("bf16" if "bf16" in fused.SA_FORWARD else "fp32")
    pk = peaks()
    B, K, W = args.batch, args.steps, max(args.warmup, 3)

    torch.manual_seed(0)
    net = randomize_bn_stats(Pointnet2Backbone(input_feature_dim=129, precision=precision)).eval().to(dev)
    sd_cpu = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    # this rank's scenes: global scene ids rank*B .. rank*B+B-1 (sharded by scene, no collective)
    host = torch.from_numpy(make_batch(B, args.points, 129, first_seed=rank * B))
    host_pool = [host.pin_memory(), host.roll(1, 0).contiguous().pin_memory()]
    pool = [h.to(dev) for h in host_pool]                    # each batch 169 MB > 126 MB L2
    in_bytes = host.numel() * 4

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # nvidia-smi is started well before the timed region: its start-up takes driver locks that stall kernel
    # launches for tens of milliseconds
    sampler = ClockSampler(local_rank) if rank == 0 else None
    with torch.no_grad():
        # ---- device-resident throughput -----------------------------------------------------
        for i in range(W):
            net({"point_clouds": pool[i % 2]})
        barrier()
        time.sleep(1.0)            # every rank: let rank 0's nvidia-smi finish starting up
        barrier()
        t_begin = time.time()
        # K steps, `lanes` of them in flight: the sampling chain of a batch is a serial, latency-bound
        # kernel on 128 SMs; batches on other streams fill the machine meanwhile
        base = torch.cuda.current_stream(dev)
        part = None
        if args.fps_sms > 0:
            from situation3d_b200.streams import SmPartition
            part = SmPartition(args.fps_sms, dev)
            net.sm_partition = part
            net._side_streams.clear()
        lanes = [part.stream(part.MAIN) if part else torch.cuda.Stream(device=dev) for _ in range(max(1, args.lanes))]
        outs = [None] * len(lanes)

        graphs = None
        if not args.no_graphs:
            # one captured step per lane (situation3d_b200.graphs): replaying it costs the host one cudaGraphLaunch
            # instead of ~0.8 ms of Python launches, which is what bounds the eager loop once lanes overlap
            from situation3d_b200.graphs import GraphedBackbone
            try:
                graphs = [GraphedBackbone(net, pool[i % 2], stream=ln, static_input=pool[i % 2]) for i, ln in enumerate(lanes)]
            except Exception as ex:      # capture unavailable (e.g. under a profiler): same kernels, enqueued from Python
                graphs, args.no_graphs = None, True
                print("bench: CUDA-graph capture failed (%s); eager launches" % repr(ex)[:120], file=sys.stderr)
                torch.cuda.synchronize()

        def run_steps(steps):
            for i in range(steps):
                ln = i % len(lanes)
                if graphs:
                    outs[ln] = graphs[ln]()
                else:
                    with torch.cuda.stream(lanes[ln]):
                        outs[ln] = net({"point_clouds": pool[i % 2]})

        for ln in lanes:
            ln.wait_stream(base)
        run_steps(max(2 * len(lanes), W))                     # untimed: every lane twice, so that the caching
        for ln in lanes:                                      # allocator owns both buffer sets a lane alternates
            base.wait_stream(ln)                              # between (a cudaMalloc inside the timed region
        torch.cuda.synchronize()                              # would serialise the device)
        for ln in lanes:
            ln.wait_stream(base)
        run_steps(len(lanes))
        for ln in lanes:
            base.wait_stream(ln)
        barrier()
        from situation3d_b200._lib import lib as _pn2
        launches0 = _pn2.pn2_launch_count()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(base)
        for ln in lanes:
            ln.wait_event(s)
        t_host0 = time.perf_counter()
        run_steps(K)
        host_ms_per_step = 1e3 * (time.perf_counter() - t_host0) / K     # host time to ENQUEUE a step (no sync inside)
        launches = int(_pn2.pn2_launch_count() - launches0)      # this library's kernels, counted at the launch sites
        if graphs:
            launches = K * graphs[0].launches_per_replay           # (counted at capture; a replay launches the same kernels)
        for ln in lanes:
            base.wait_stream(ln)
        e.record(base)
        torch.cuda.synchronize()
        dt = reduce_max(s.elapsed_time(e) * 1e-3)
        out = outs[0]
        barrier()

        # ---- end to end: pinned host input -> H2D -> forward -> D2H of the result, double-buffered ----
        # each lane owns a device input buffer and pinned result buffers; its H2D copy, forward and D2H
        # copies are enqueued on the lane's stream, so lanes overlap each other's copies and kernels
        # ---- end to end through the public serving API (situation3d_b200.graphs.BackbonePipeline): per step a pinned
        # host batch is copied to the lane's device buffer, the captured step is replayed and fp2_features / fp2_xyz /
        # fp2_inds are copied back to pinned host memory, all on the lane's stream; `lanes` batches in flight
        pipe, dbuf, res_host = None, None, None
        if not args.no_e2e:
            graphs = None                                     # release the resident-run graphs' pools
            if not args.no_graphs:
                from situation3d_b200.graphs import BackbonePipeline
                pipe = BackbonePipeline(net, pool[0], streams=lanes)
                in_bytes, out_bytes = pipe.h2d_bytes, pipe.d2h_bytes
            else:
                dbuf = [pool[0].clone() for _ in lanes]
                res_host = [{k: torch.empty_like(out[k], device="cpu").pin_memory() for k in ("fp2_features", "fp2_xyz", "fp2_inds")}
                            for _ in lanes]
                out_bytes = sum(v.numel() * v.element_size() for v in res_host[0].values())
        else:
            out_bytes = 0

        def e2e_loop(steps):
            for i in range(steps):
                if pipe is not None:
                    pipe.submit(host_pool[i % 2])
                    continue
                ln = i % len(lanes)
                with torch.cuda.stream(lanes[ln]):
                    dbuf[ln].copy_(host_pool[i % 2], non_blocking=True)
                    o = net({"point_clouds": dbuf[ln]})
                    for k, v in res_host[ln].items():
                        v.copy_(o[k], non_blocking=True)
                    outs[ln] = o

        dt_e2e = float("nan")
        if not args.no_e2e:
            for ln in lanes:
                ln.wait_stream(base)
            e2e_loop(max(2 * len(lanes), W))
            for ln in lanes:
                base.wait_stream(ln)
            barrier()
            s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s2.record(base)
            for ln in lanes:
                ln.wait_event(s2)
            e2e_loop(K)
            for ln in lanes:
                base.wait_stream(ln)
            e2.record(base)
            torch.cuda.synchronize()
            dt_e2e = reduce_max(s2.elapsed_time(e2) * 1e-3)
        clocks = sampler.stop(t_begin, time.time()) if sampler else None
        barrier()

        rows, ref_cuda, cpu_base = None, None, None
        if world > 1:
            # the per-kernel breakdown, the CPU baseline and the reference-CUDA arm are single-GPU measurements
            args.no_kernel_breakdown = args.no_reference_cuda = args.no_cpu_baseline = True
        if rank == 0:
            if not args.no_kernel_breakdown:
                rows = kernel_breakdown(net, pool[0], precision, pk)
            try:
                ref_cuda = None if args.no_reference_cuda else reference_cuda_arm(pool[0], sd_cpu)
            except Exception as ex:   # test infrastructure must not take the bench down
                ref_cuda = {"unavailable": repr(ex)[:200]}
            if not args.no_cpu_baseline:
                if prev_affinity:
                    os.sched_setaffinity(0, prev_affinity)      # the CPU baseline gets every host core again
                from oracle import pn2_oracle as orc
                cores = os.cpu_count() or 1
                torch.set_num_threads(cores)
                sample = host[: min(B, 4)].contiguous()
                orc.backbone(sample[:1].contiguous(), sd_cpu)
                t0 = time.perf_counter()
                orc.backbone(sample, sd_cpu)
                t1 = time.perf_counter() - t0
                cpu_base = {"value": sample.shape[0] / t1, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "%d of the step's %d scenes, one pass after a 1-scene warm-up; oracle port "
                                      "(PyTorch CPU wiring over oracle/pn2_oracle.c, OpenMP + %d torch threads)"
                                      % (sample.shape[0], B, cores)}
        barrier()

    if rank == 0:
        line = {"metric": METRIC, "value": world * B * K / dt, "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": "Pointnet2Backbone forward (SA1-SA4 + FP1-FP2), B=%d scenes/GPU x %d points, "
                                       "xyz+height+128-d multiview (BASELINE configs[1])" % (B, args.points),
                           "scenes_per_gpu": B, "points": args.points, "feature_channels": 129,
                           "precision": precision, "sharding": "by scene, no collective",
                           "batches_in_flight": len(lanes), "cuda_graphs": not args.no_graphs,
                           "sm_partition": {"fps": part.sms[0], "main": part.sms[1]} if part else None,
                           "l2": "each input batch is %.0f MB (> 126 MB L2); two batches alternate" % (in_bytes / 1e6)},
                "e2e": {"value": world * B * K / dt_e2e, "unit": UNIT, "ms_per_step": 1e3 * dt_e2e / K,
                        "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                        "api": "situation3d_b200.graphs.BackbonePipeline.submit(pinned host point_clouds) -> pinned host results"
                               if not args.no_graphs else "Pointnet2Backbone.forward(data_dict) on pinned host point_clouds",
                        "host_numa_binding": "rank pinned to its GPU's %d local CPUs before allocating pinned buffers" % numa_cpus
                                             if numa_cpus else None},
                "gpu_launches": launches * world, "gpu_launches_per_step": launches / K,
                "host_enqueue_ms_per_step": host_ms_per_step, "clocks": clocks, "peaks": pk}
        if rows:
            dom = max(rows, key=lambda r: r["share"])
            traffic = ncu_traffic()
            line["roofline"] = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": dom["achieved"],
                                "peak": dom["peak"], "unit": dom["unit"], "frac": dom["frac"],
                                "traffic": traffic.get(dom["kernel"]),
                                "share_of_step": dom["share"], "us_per_launch": dom["us"],
                                "peak_source": pk["source"] + (" burst" if dom["bound"] == "tensor" else "")}
            for extra in ("rounds_per_s", "point_updates_per_s", "fp32_issue_frac"):
                if extra in dom:
                    line["roofline"][extra] = dom[extra]
            if "rounds_per_s" in dom:
                line["roofline"]["note"] = ("latency chain of dependent argmax rounds: neither HBM nor the tensor pipe bounds it; "
                                            "rounds_per_s is the figure of merit (DESIGN.md 4.1)")
            line["roofline_kernels"] = [
                {k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items()
                 if k in ("kernel", "bound", "us", "share", "achieved", "unit", "frac", "hbm_gbs", "rounds_per_s",
                          "fp32_issue_frac", "layer_tflops")}
                for r in rows]
            for r in line["roofline_kernels"]:
                if r["kernel"] in traffic:
                    r["traffic"] = traffic[r["kernel"]]
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        if ref_cuda:
            line["reference_cuda"] = ref_cuda
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517")] + sys.argv
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
