#!/usr/bin/env python
"""bench.py -- Pointnet2Backbone forward throughput (scenes/s) on synthetic ScanNet-shaped scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload = BASELINE.json configs[1] (--config 2, the default): B=8 scenes per GPU, 40 000 points, xyz + height
+ 128-d multiview features, SA1-SA4 + FP1-FP2, bf16 fused MLPs.  A step = one backbone forward over one batch.
Scenes are sharded by rank with no collective on the data path (weak scaling).  The other configurations of
BASELINE.json run as --config 1 / 3 / 5 and, on the default run, as short sub-records under "configs":
  1  one scene, fp32 arm, one batch in flight (latency)
  3  backbone + situation re-encoding of the first 256 seeds, 4 scenes per GPU (32 scenes over 8 GPUs)
  5  stress: 100k / 150k / 200k-point scenes, SA1 npoint 4096 / nsample 64, 8 scenes per GPU (64 over 8 GPUs)

One JSON line on rank 0:
  value     whole-job scenes/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the public serving API (situation3d_b200.graphs.BackbonePipeline) with HOST (pinned)
            inputs: H2D of the batch and D2H of fp2_features / fp2_xyz / fp2_inds inside the timed region.  Input
            format: fp32 coordinates + bf16 feature rows (Pointnet2Backbone.pack_point_clouds; bit-identical results,
            half the bytes); "e2e_f32_input" is the same loop fed the reference's fp32 point_clouds
  parity    one scene of a TIMED batch against the CPU oracle: every index tensor bit for bit, fp2_features by
            oracle/parity.py; the process exits non-zero when the gate fails
  roofline  the dominant kernel of the step (largest share of device time), timed live with CUDA events;
            roofline_kernels lists every kernel of the step the same way; roofline_step = the whole step against the
            HBM and tensor ceilings of SURVEY.md 8d
  cpu_baseline    the CPU oracle (oracle/) on this box's host cores, bounded sample, rank 0 only
  reference_cuda  the reference's own CUDA kernels + its own unmodified Python modules (oracle/_ref: extension and
            byte code compiled from /root/reference) + stock PyTorch conv / BN / max-pool on the same GPU -- the
            existing GPU implementation; also compared with this library's outputs on the same batch

--impl reference times the reference's path on the host CPU (the reference ships no CPU kernels, its ops assert
"CPU not supported", so this is the oracle port: same wiring in PyTorch CPU ops over the C restatement of its
kernels, all host threads); it never imports the product package.
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

# BASELINE.json configs: per-GPU batch, points, precision, batches in flight, SA pyramid.
# Batches in flight ("lanes", one CUDA stream + graph each) are what fills the GPU: a batch's sampling chains occupy
# 8 CTAs per scene (one per scene above 81 920 points) for milliseconds.  Measured optima (scripts/gpu_r2_check7.sh):
# config 2 plateaus at 16-24 (12.8-13.0 k scenes/s; 11.4-12.0 k at 8), config 3 at 24, config 5 at 16-24 (3.3 k against
# 1.0 k scenes/s at 2).
CONFIGS = {
    1: dict(batch=1, points=40000, precision="fp32", lanes=1, npoints=(2048, 1024, 512, 256),
            what="config 1: Pointnet2Backbone forward, one 40 000-point scene, fp32 arm, one batch in flight"),
    2: dict(batch=8, points=40000, precision="bf16", lanes=20, npoints=(2048, 1024, 512, 256),
            what="config 2: Pointnet2Backbone forward (SA1-SA4 + FP1-FP2), B=8 scenes/GPU x 40000 points, "
                 "xyz+height+128-d multiview, bf16 fused MLPs (BASELINE configs[1])"),
    3: dict(batch=4, points=40000, precision="bf16", lanes=24, npoints=(2048, 1024, 512, 256), reencode=True,
            what="config 3: backbone + situation re-encoding (agent-frame transform + pos_embed + prior) of the first 256 "
                 "seeds, 4 scenes/GPU (32 scenes sharded over 8 GPUs)"),
    5: dict(batch=8, points=200000, precision="bf16", lanes=24, npoints=(4096, 2048, 1024, 512),
            what="config 5: stress, 200 000-point scenes, SA npoint 4096/2048/1024/512, nsample 64/32/16/16, "
                 "8 scenes/GPU (64 scenes over 8 GPUs)"),
}
CONFIGS[4] = dict(batch=8, points=40000, precision="fp32", lanes=1, npoints=(2048, 1024, 512, 256), train=True,
                  what="config 4: SQA3D-shaped training step, backbone forward + backward (train-mode BatchNorm, loss = "
                       "mean(fp2_features^2)), SGD, gradient all-reduce over NCCL overlapped with backward, 8 scenes/GPU")
RADII, NSAMPLES = (0.2, 0.4, 0.8, 1.2), (64, 32, 16, 16)
INDEX_KEYS = ("sa1_inds", "sa2_inds", "sa3_inds", "sa4_inds", "fp2_inds")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--precision", default=None, help="override the configuration's arm: fp32 | bf16")
    ap.add_argument("--batch", type=int, default=None, help="override scenes per GPU per step")
    ap.add_argument("--points", type=int, default=None, help="override points per scene")
    ap.add_argument("--lanes", type=int, default=None,
                    help="batches in flight per GPU: step i runs on CUDA stream i %% lanes (1 = strictly serial steps)")
    ap.add_argument("--fps-sms", type=int, default=0,
                    help="give the sampling chains their own group of >= this many SMs (CUDA green contexts); 0 = off")
    ap.add_argument("--sampling-mode", default="auto", choices=["auto", "latency", "throughput"],
                    help="FPS kernel of the captured steps for mid-sized scenes (auto: throughput from 128 scenes in flight)")
    ap.add_argument("--no-graphs", action="store_true",
                    help="enqueue every step from Python instead of replaying a captured CUDA graph per lane")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-breakdown", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--no-sub-configs", action="store_true", help="skip the short runs of the other configurations")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end loop")
    ap.add_argument("--no-parity", action="store_true", help="profiling runs only: skip the oracle gate")
    return ap.parse_args()


def config_of(args, which=None):
    cfg = dict(CONFIGS[which if which is not None else args.config])
    if which is None:
        for k in ("precision", "batch", "points", "lanes"):
            if getattr(args, k) is not None:
                cfg[k] = getattr(args, k)
    return cfg


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


def oracle_layers(cfg):
    return tuple(("sa%d" % (i + 1), cfg["npoints"][i], RADII[i], NSAMPLES[i]) for i in range(4))


# ---------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's path on the host CPU (oracle port), the configuration's full batch per step.  Imports nothing
    from the product package: scenes and weights come from oracle/workload.py."""
    if rank != 0:
        return
    import torch
    from oracle import pn2_oracle as orc
    from oracle import workload
    cfg = config_of(args)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = workload.backbone_state_dict(129, seed=0)
    scenes = cfg["batch"]                            # the step's whole batch (8 scenes for the headline configuration)
    pc = torch.from_numpy(workload.synthetic().make_batch(scenes, cfg["points"], 129))
    layers = oracle_layers(cfg)
    with torch.no_grad():
        t0 = time.perf_counter()
        orc.backbone(pc, sd, layers)                 # first call: page-in + thread pools (counts as the first warm-up step)
        first = time.perf_counter() - t0
        # bounded: the whole run (warm-up + timed steps) stays within ~4 minutes of CPU time
        budget = 240.0
        warm = max(0, min(args.warmup - 1, int(0.2 * budget / max(first, 1e-3))))
        for _ in range(warm):
            orc.backbone(pc, sd, layers)
        steps = max(1, min(args.steps, int(0.8 * budget / max(first, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.backbone(pc, sd, layers)
        dt = time.perf_counter() - t0
    value = scenes * steps / dt
    sample = "%d scenes of %d points per step (the configuration's whole per-GPU batch) x %d steps; oracle port: PyTorch " \
             "CPU wiring over oracle/pn2_oracle.c (OpenMP), %d torch threads" % (scenes, cfg["points"], steps, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm + 1, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["what"], "scenes_per_step": scenes, "points": cfg["points"],
                       "feature_channels": 129, "device": "host CPU"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def kernel_breakdown(net, pc, precision, pk, reps=5, sampling_mode="latency"):
    """Every kernel of one step launched alone through the C ABI and timed with CUDA events on the
    launching (current) stream; algorithmic bytes / FLOPs per launch as defined in DESIGN.md.  ``sampling_mode``: the
    FPS kernel the timed step used (fused.sampling_mode), so that the ranking describes that step."""
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
        with fused.sampling_mode(sampling_mode):
            inds, cxyz = fused.fps_with_xyz(src_xyz, m.npoint)
            t = timeit(lambda: fused.fps_with_xyz(src_xyz, m.npoint))
            bucketed = fused.lib.pn2_furthest_point_sampling_workspace_bytes_mode(B, n_in, m.npoint, fused._SAMPLING_MODE[0]) > 0
        # FPS is a chain of dependent rounds: besides the (meaningless) HBM figure, report rounds/s and the share of the
        # machine the launch occupies (one CTA per scene for large scenes, csrc/fps_bucket.cu)
        # SMs a sampling launch can occupy: the cluster kernel (csrc/fps.cu) runs 8 CTAs of 128 threads per scene for
        # 8 193 .. 81 920 points, three per SM (two for a single scene: the all-register variant); one CTA per scene otherwise
        # (small scenes: 256-512 threads, about half an SM; csrc/fps_bucket.cu beyond)
        if bucketed:
            sm_share = min(1.0, B / 148.0)                     # csrc/fps_bucket.cu: one CTA per scene, one per SM
        elif 8192 < n_in <= 81920:
            sm_share = min(1.0, B * 8 / (3.0 if B > 1 else 2.0) / 148.0)
        else:
            sm_share = min(1.0, B * (0.5 if n_in <= 8192 else 1.0) / 148.0)
        rows.append({"kernel": "fps_" + name, "variant": "bucketed, one CTA per scene" if bucketed else "register-resident",
                     "bound": "hbm", "seconds": t,
                     "alg_bytes": B * (12 * n_in + 16 * m.npoint), "rounds_per_s": (m.npoint - 1) / t,
                     "sm_share": sm_share})
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
    from situation3d_b200._lib import lib as _pn2
    for name, (u, k), skip_rows in (("fp1", (2, 3), feats_rows[2]), ("fp2", (1, 2), feats_rows[1])):
        un, kn = cxyz_all[u], cxyz_all[k]
        img = imgs[4 if name == "fp1" else 5]
        dims = img.dims
        n = un.shape[1]
        flops = 2 * B * n * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
        nn_bytes = B * (12 * (un.shape[1] + kn.shape[1]) + 24 * un.shape[1])
        abytes = B * (4 * known_rows.shape[2] * kn.shape[1] + 4 * skip_rows.shape[2] * n + 24 * n + 2 * 4 * dims[-1] * n) \
            + 4 * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
        whole = lambda: fused.fp_layer(precision, img, un, kn, known_rows, skip_rows)
        n0 = _pn2.pn2_launch_count()
        _, out_rows = whole()
        if _pn2.pn2_launch_count() - n0 == 1:
            # three_nn + interpolation + concat + MLP in one launch (csrc/fp_tc2.cu)
            t = timeit(whole)
            rows.append({"kernel": "%s_layer_%s" % (name, precision), "bound": "tensor", "seconds": t,
                         "alg_flops": flops, "alg_bytes": abytes + nn_bytes - 24 * B * n})
        else:
            d2, i3 = fused.three_nn(un, kn)
            t = timeit(lambda: fused.three_nn(un, kn))
            rows.append({"kernel": "three_nn_" + name, "bound": "hbm", "seconds": t, "alg_bytes": nn_bytes})
            run = lambda: fused.FP_FORWARD[precision](img, d2, i3, known_rows, skip_rows)
            _, out_rows = run()
            t = timeit(run)
            rows.append({"kernel": "%s_fused_%s" % (name, precision), "bound": "tensor", "seconds": t,
                         "alg_flops": flops, "alg_bytes": abytes})
        known_rows = out_rows
    # share of the step: a kernel's time weighted by the fraction of the SMs its launch can occupy (the sampling kernels
    # run one CTA per scene; with several batches in flight the rest of the machine runs other batches meanwhile)
    total = sum(r["seconds"] for r in rows)
    total_sm = sum(r["seconds"] * r.get("sm_share", 1.0) for r in rows)
    for r in rows:
        r["share_serial"] = r["seconds"] / total
        r["share"] = r["seconds"] * r.get("sm_share", 1.0) / total_sm
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


def reference_cuda_arm(pc, sd, ours, steps=24, lanes=8):
    """The existing GPU implementation on the same B200: the reference's own CUDA kernels (oracle/_ref/pn2_ref_ext.so,
    unmodified sources) driven by the reference's own unmodified Python modules (oracle/_ref/*.bytecode, byte code of
    lib/pointnet2/{pointnet2_modules,pointnet2_utils,pytorch_utils}.py) with stock PyTorch conv / BN / ReLU / max-pool
    (cuDNN, TF32 allowed as by torch's default), composed as Pointnet2Backbone (oracle/ref_modules.py), `lanes` batches
    in flight on separate streams like this library's arm.  Also compares its outputs with `ours` on the same batch."""
    import torch
    from oracle import ref_modules
    from oracle.build_ref import load_ref_ext
    ext = load_ref_ext()
    if ext is None or not ref_modules.available():
        return None
    rm = ref_modules.load(ext)
    net = ref_modules.make_backbone(rm, pc.shape[2] - 3).eval()
    net.load_state_dict(sd)
    net = net.to(pc.device)
    streams = [torch.cuda.Stream(device=pc.device) for _ in range(lanes)]
    base = torch.cuda.current_stream(pc.device)
    with torch.no_grad():
        out = net(pc)
        for st in streams:                                   # warm every stream's allocator pool
            st.wait_stream(base)
            with torch.cuda.stream(st):
                net(pc)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(base)
        for st in streams:
            st.wait_event(s)
        for i in range(steps):
            with torch.cuda.stream(streams[i % lanes]):
                net(pc)
        for st in streams:
            base.wait_stream(st)
        e.record(base)
        e.synchronize()
    dt = s.elapsed_time(e) * 1e-3
    rec = {"value": pc.shape[0] * steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / steps, "steps": steps,
           "batches_in_flight": lanes,
           "what": "reference _ext_src kernels (compiled unmodified for sm_100) + the reference's unmodified "
                   "pointnet2_modules / pointnet2_utils / pytorch_utils (byte code) + stock torch 1x1 conv/BN/ReLU/"
                   "max_pool2d (cuDNN, TF32 conv default), composed as Pointnet2Backbone, fp32, B=%d, 1 GPU" % pc.shape[0]}
    if ours is not None:
        from oracle.parity import feature_error
        agree = {k: bool(torch.equal(ours[k], out[k])) for k in INDEX_KEYS}
        st = feature_error(ours["fp2_features"], out["fp2_features"])
        rec["ours_vs_reference_cuda"] = {"indices_bit_exact": agree, "fp2_features_rel_l2": st["rel_l2"],
                                         "fp2_features_max_err_over_max_ref": st["max_err"] / max(st["max_ref"], 1e-30),
                                         "scenes_compared": int(pc.shape[0])}
    return rec


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


def parity_gate(out, host_batch, sd_cpu, cfg, precision, situations=None, reenc=None, scenes=1):
    """`scenes` scenes of a timed batch against the CPU oracle: index tensors bit for bit, fp2_features (and the
    re-encoded tokens of config 3) by oracle/parity.py.  Returns a JSON-able record with "ok"."""
    import torch
    from oracle import pn2_oracle as orc
    from oracle.parity import check_features, feature_error
    t0 = time.perf_counter()
    with torch.no_grad():
        want = orc.backbone(host_batch[:scenes].contiguous(), sd_cpu, oracle_layers(cfg))
    rec = {"scenes": scenes, "against": "CPU oracle (oracle/pn2_oracle.c + PyTorch CPU wiring), same weights and inputs",
           "indices_bit_exact": {}}
    ok = True
    for k in INDEX_KEYS:
        same = bool(torch.equal(out[k][:scenes].cpu(), want[k]))
        rec["indices_bit_exact"][k] = same
        ok &= same
    same = bool(torch.equal(out["fp2_xyz"][:scenes].cpu(), want["fp2_xyz"]))
    rec["fp2_xyz_bit_exact"] = same
    ok &= same
    got = out["fp2_features"][:scenes].float().cpu()
    fok, msg = check_features(got, want["fp2_features"], precision)
    st = feature_error(got, want["fp2_features"])
    rec["fp2_features"] = {"ok": fok, "gate": "per tensor: max|err| <= 2e-2 max|ref| and rel L2 <= 1e-2" if precision == "bf16"
                           else "every element: |err| <= 1e-3 max(|ref|, rms(ref))",
                           "max_err": st["max_err"], "max_ref": st["max_ref"], "rms_ref": st["rms_ref"],
                           "rel_l2": st["rel_l2"], "elementwise_frac_within_2e-2": st["frac_within_2e-2"],
                           "elementwise_frac_within_1e-3": st["frac_within_1e-3"]}
    ok &= fok
    if reenc is not None:
        pe = reenc.pos_embed
        t = out["att_feat_pre"].shape[1]
        tok = want["fp2_features"][:, :, :t].transpose(1, 2).contiguous()
        want_tok, want_pos, want_prior = orc.reencode(tok, want["fp2_xyz"][:, :t].contiguous(), situations[:scenes].cpu(),
                                                      pe[0].weight.detach().cpu(), pe[0].bias.detach().cpu(),
                                                      pe[2].weight.detach().cpu(), pe[2].bias.detach().cpu())
        rok, rmsg = check_features(out["scene_feat"][:scenes].cpu(), want_tok, precision)
        pos_ok = bool(torch.allclose(out["scene_positions_agent"][:scenes].cpu(), want_pos, rtol=1e-5, atol=1e-6))
        prior_ok = bool(torch.allclose(out["auxiliary_task_loc_gt"][:scenes].cpu(), want_prior, rtol=1e-3, atol=1e-7))
        rec["reencode"] = {"tokens_ok": rok, "tokens": rmsg, "positions_ok": pos_ok, "prior_ok": prior_ok}
        ok &= rok and pos_ok and prior_ok
    rec["ok"] = bool(ok)
    rec["oracle_seconds"] = round(time.perf_counter() - t0, 2)
    return rec


class Runner:
    """One configuration on this rank's GPU: resident throughput (+ optional end-to-end loops) and the parity gate."""

    def __init__(self, args, cfg, rank, local_rank, world):
        import torch
        self.args, self.cfg, self.rank, self.world = args, cfg, rank, world
        self.dev = torch.device("cuda", local_rank)
        self.B, self.N = cfg["batch"], cfg["points"]
        self.precision = cfg["precision"]

    def build(self):
        import torch
        from situation3d_b200.backbone_module import Pointnet2Backbone
        from situation3d_b200.synthetic import make_batch, make_situations, randomize_bn_stats
        cfg, dev = self.cfg, self.dev
        torch.manual_seed(0)
        net = randomize_bn_stats(Pointnet2Backbone(input_feature_dim=129, precision=self.precision,
                                                   npoints=cfg["npoints"])).eval().to(dev)
        self.backbone = net
        self.sd_cpu = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
        self.reenc, self.situations = None, None
        self.model = net
        if cfg.get("reencode"):
            from situation3d_b200.scene_encoder import SituatedSceneEncoder
            enc = SituatedSceneEncoder(129, 256, precision=self.precision).eval().to(dev)
            enc.backbone_net = net
            self.model, self.reenc = enc, enc.reencoder
            self.situations = torch.from_numpy(make_situations(self.B, seed=self.rank)).to(dev)
        # this rank's scenes: global scene ids rank*B .. rank*B+B-1 (sharded by scene, no collective)
        self.host = torch.from_numpy(make_batch(self.B, self.N, 129, first_seed=self.rank * self.B))
        # two batches alternate (the second = the first with its scenes rotated); each is larger than L2 for B >= 6
        self.host_pool = [self.host.pin_memory(), self.host.roll(1, 0).contiguous().pin_memory()]
        self.pool = [h.to(dev) for h in self.host_pool]
        return self

    def inputs(self, pc):
        d = {"point_clouds": pc}
        if self.situations is not None:
            d["auxiliary_task"] = self.situations
        return d

    def barrier(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.MIN)
        return float(t.item())

    def resident(self, K, W, part=None):
        """K timed steps, `lanes` batches in flight, inputs resident in HBM.  Returns (seconds, launches, host ms/step)."""
        import torch
        from situation3d_b200._lib import lib as _pn2
        from situation3d_b200.streams import lane_stream
        args, dev, model = self.args, self.dev, self.model
        base = torch.cuda.current_stream(dev)
        nl = max(1, self.cfg["lanes"])
        self.lanes = [part.stream(part.MAIN) if part else lane_stream(dev) for _ in range(nl)]
        lanes = self.lanes
        outs = [None] * nl
        with torch.no_grad():
            for i in range(W):
                model(self.inputs(self.pool[i % 2]))
            self.barrier()
            self.graphs = None
            if not args.no_graphs:
                # one captured step per lane (situation3d_b200.graphs): replaying it costs the host one cudaGraphLaunch
                # instead of ~0.8 ms of Python launches, which is what bounds the eager loop once lanes overlap
                from situation3d_b200.graphs import GraphedBackbone, pick_sampling_mode
                # which FPS kernel the captured steps use at 40 000 points: the fastest launch, or -- with enough scenes in
                # flight to give every SM a scene -- the one with the least SM time (identical results; fused.sampling_mode)
                self.sampling_mode = (args.sampling_mode if args.sampling_mode != "auto"
                                      else pick_sampling_mode(nl * self.cfg["batch"]))
                try:
                    self.graphs = [GraphedBackbone(model, self.inputs(self.pool[i % 2]), stream=ln,
                                                   static_input=self.inputs(self.pool[i % 2]),
                                                   sampling_mode=self.sampling_mode) for i, ln in enumerate(lanes)]
                except Exception as ex:      # capture unavailable (e.g. under a profiler): same kernels, enqueued from Python
                    self.graphs, args.no_graphs = None, True
                    print("bench: CUDA-graph capture failed (%s); eager launches" % repr(ex)[:120], file=sys.stderr)
                    torch.cuda.synchronize()
            graphs = self.graphs

            def run_steps(steps):
                for i in range(steps):
                    ln = i % nl
                    if graphs:
                        outs[ln] = graphs[ln]()
                    else:
                        with torch.cuda.stream(lanes[ln]):
                            outs[ln] = model(self.inputs(self.pool[ln % 2]))

            for ln in lanes:
                ln.wait_stream(base)
            run_steps(max(2 * nl, W))                             # untimed: every lane twice, so that the caching
            for ln in lanes:                                      # allocator owns both buffer sets a lane alternates
                base.wait_stream(ln)                              # between (a cudaMalloc inside the timed region
            torch.cuda.synchronize()                              # would serialise the device)
            for ln in lanes:
                ln.wait_stream(base)
            run_steps(nl)
            for ln in lanes:
                base.wait_stream(ln)
            self.barrier()
            launches0 = _pn2.pn2_launch_count()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(base)
            for ln in lanes:
                ln.wait_event(s)
            t_host0 = time.perf_counter()
            run_steps(K)
            host_ms = 1e3 * (time.perf_counter() - t_host0) / K          # host time to ENQUEUE a step (no sync inside)
            launches = int(_pn2.pn2_launch_count() - launches0)          # this library's kernels, counted at the launch sites
            if graphs:
                launches = K * graphs[0].launches_per_replay             # (counted at capture; a replay launches the same kernels)
            for ln in lanes:
                base.wait_stream(ln)
            e.record(base)
            torch.cuda.synchronize()
            dt = self.reduce(s.elapsed_time(e) * 1e-3)
            # lane 0's last TIMED batch (its input is pool[0] = self.host), copied out of the graph's memory pool
            self.out0 = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in outs[0].items() if k != "point_clouds"}
            self.barrier()
        return dt, launches, host_ms

    def e2e(self, K, W, compact):
        """Pinned host batch -> H2D -> captured step -> D2H of the results, per lane, through BackbonePipeline.
        Returns (seconds, h2d bytes, d2h bytes, last result of lane 0 as host tensors)."""
        import torch
        from situation3d_b200.graphs import BackbonePipeline
        dev, lanes = self.dev, self.lanes
        base = torch.cuda.current_stream(dev)
        with torch.no_grad():
            if compact:
                host_in = [tuple(t.pin_memory() for t in self.backbone.pack_point_clouds(h)) for h in self.host_pool]
                example = tuple(t.to(dev) for t in host_in[0])
            else:
                host_in, example = self.host_pool, self.pool[0]
            pipe = BackbonePipeline(self.backbone, example, streams=lanes)

            def loop(steps):
                t = None
                for i in range(steps):
                    t = pipe.submit(host_in[(i % len(lanes)) % 2])
                return t

            for ln in lanes:
                ln.wait_stream(base)
            loop(max(2 * len(lanes), W))
            for ln in lanes:
                base.wait_stream(ln)
            self.barrier()
            s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s2.record(base)
            for ln in lanes:
                ln.wait_event(s2)
            loop(K)
            for ln in lanes:
                base.wait_stream(ln)
            e2.record(base)
            torch.cuda.synchronize()
            dt = self.reduce(s2.elapsed_time(e2) * 1e-3)
            res0 = {k: v.clone() for k, v in pipe.host[0].items()}     # lane 0 always carries host_in[0] = self.host
            pipe.drain()
            h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
            del pipe
        self.barrier()
        return dt, h2d, d2h, res0

    def gate(self, out=None, scenes=1):
        out = out if out is not None else self.out0
        rec = parity_gate(out, self.host, self.sd_cpu, self.cfg, self.precision, self.situations, self.reenc, scenes)
        rec["all_ranks_ok"] = bool(self.reduce(1.0 if rec["ok"] else 0.0, "min") > 0.5)
        return rec


def train_config(args, rank, local_rank, world, K):
    """BASELINE.json config 4: the training step (situation3d_b200.train_step.BackboneTrainer) at 8 scenes per GPU."""
    import copy
    import torch
    import torch.distributed as dist
    from situation3d_b200.backbone_module import Pointnet2Backbone
    from situation3d_b200.synthetic import make_batch
    from situation3d_b200.train_step import BackboneTrainer
    cfg = config_of(args, 4)
    dev = torch.device("cuda", local_rank)
    B = cfg["batch"]
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True            # the reference's cuDNN convolutions run TF32 by torch's default
    torch.manual_seed(0)
    net = Pointnet2Backbone(input_feature_dim=129).to(dev)
    # parity of the row-layout training path against the reference wiring (same parameters, 2 scenes, strict fp32)
    par = None
    if not args.no_parity:
        torch.backends.cuda.matmul.allow_tf32 = False
        cudnn_tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False              # strict fp32 on BOTH sides (the reference wiring convolves with cuDNN)
        a, b = copy.deepcopy(net).train(), copy.deepcopy(net).train()
        b.train_layout = "reference"
        pc2 = torch.from_numpy(make_batch(2, cfg["points"], 129, first_seed=rank * B)).to(dev)
        la = a({"point_clouds": pc2})["fp2_features"].square().mean()
        lb = b({"point_clouds": pc2})["fp2_features"].square().mean()
        la.backward()
        lb.backward()
        worst = max(float((p1.grad - p2.grad).abs().max()) / (float(p2.grad.abs().max()) + 1e-12)
                    for p1, p2 in zip(a.parameters(), b.parameters()))
        worst_l2 = max(float((p1.grad - p2.grad).norm() / (p2.grad.norm() + 1e-30)) for p1, p2 in zip(a.parameters(), b.parameters()))
        rel = abs(float(la.detach()) - float(lb.detach())) / max(abs(float(lb.detach())), 1e-30)
        # Gate: loss to 1e-4; every parameter's gradient to 2e-2 in relative L2 and 1e-1 of its largest entry.  The forward
        # pass agrees to 1e-5; the gradients differ by a few 1e-3 because max-pool argmax / ReLU decisions between nearly
        # equal values flip with the last bit of the BatchNorm arithmetic (x*a + b here, (x-mean)*invstd*w + b in
        # torch): the SAME deviations, digit for digit, appear when this path's BatchNorm is written as x*a + b in plain
        # torch ops (PN2_TRAIN_FUSED_BN=2, scripts/train_parity_diag.py, profiles/r2_train_parity_diag.txt), so they
        # measure rounding sensitivity, not the kernels.  The reference wiring differs from itself by 1e-5 (atomics).
        par = {"against": "the reference's operator-by-operator wiring (QueryAndGroup -> NCHW SharedMLP -> max_pool2d) with "
                          "autograd through the *_grad kernels, same parameters, 2 scenes, strict fp32 on both sides (no TF32)",
               "loss_rel_diff": rel, "max_grad_diff_over_max_grad": worst, "max_grad_rel_l2": worst_l2,
               "gate": "loss <= 1e-4, per-parameter gradient rel L2 <= 2e-2 and max |diff| <= 1e-1 max |grad|",
               "ok": bool(rel <= 1e-4 and worst_l2 <= 2e-2 and worst <= 1e-1)}
        del a, b, pc2
        torch.cuda.empty_cache()
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = cudnn_tf32
    tr = BackboneTrainer(net)
    # two alternating batches, as a data loader would hand them over: step i is given batch i+1 so that its sampling
    # chain (coordinates only) runs under step i's backward pass (BackboneTrainer.step(next_point_clouds=...))
    pcs = [torch.from_numpy(make_batch(B, cfg["points"], 129, first_seed=rank * B + 100 * j)).to(dev) for j in range(2)]
    pc = pcs[0]

    def timed(nsteps, prefetch):
        for i in range(3):
            tr.step(pcs[i % 2], pcs[(i + 1) % 2] if prefetch else None)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps + 1)]
        evs[0].record()
        for i in range(nsteps):
            l_ = tr.step(pcs[(i + 1) % 2], pcs[i % 2] if prefetch else None)
            evs[i + 1].record()
        torch.cuda.synchronize()
        d_ = evs[0].elapsed_time(evs[nsteps]) * 1e-3
        # the step's host enqueue time (~8 ms of Python / autograd per step) is close to its GPU time, so one host hiccup
        # shows in the mean; the median of the per-step times is reported beside it
        med_ = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(nsteps))[nsteps // 2]
        if world > 1:
            t_ = torch.tensor([d_, med_], dtype=torch.float64, device=dev)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            d_, med_ = float(t_[0].item()), float(t_[1].item())
        return d_, l_, med_

    dt_np, _, med_np = timed(max(2, K // 2), False)
    dt, loss, med = timed(K, True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the reference's operator-by-operator wiring of the same step (this library's kernels behind its nine ops, NCHW
    # SharedMLP by cuDNN, max_pool2d), same parameters: the "existing implementation" bar of this configuration
    ref_ms = None
    if rank == 0 and not args.no_reference_cuda:
        try:
            ref_net = copy.deepcopy(net)
            ref_net.train_layout = "reference"
            rt = BackboneTrainer(ref_net) if world == 1 else None
            if rt is not None:
                for _ in range(2):
                    rt.step(pc)
                torch.cuda.synchronize()
                s.record()
                for _ in range(5):
                    rt.step(pc)
                e.record()
                torch.cuda.synchronize()
                ref_ms = s.elapsed_time(e) / 5
            del ref_net, rt
        except Exception as ex:  # out of memory on a shared box must not fail the record
            ref_ms = "failed: %s" % type(ex).__name__
        torch.cuda.empty_cache()
    ar_us = None
    if world > 1:
        t = torch.zeros_like(tr.bucket.flat)
        dist.all_reduce(t)
        torch.cuda.synchronize()
        s.record()
        for _ in range(20):
            dist.all_reduce(t)
        e.record()
        torch.cuda.synchronize()
        ar_us = 1e3 * s.elapsed_time(e) / 20
    torch.backends.cuda.matmul.allow_tf32 = tf32
    rec = {"config": 4, "workload": cfg["what"], "scenes_per_gpu": B, "points": cfg["points"], "steps": K,
           "value": world * B * K / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / K, "loss": float(loss),
           "ms_per_step_median": med, "ms_per_step_without_sampling_prefetch": 1e3 * dt_np / max(2, K // 2),
           "ms_per_step_without_sampling_prefetch_median": med_np,
           "host_bound_note": "Python/autograd enqueue of the ~350 launches of a step takes about as long as the GPU needs for "
                              "them (scripts/train_prefetch_diag.py): the mean is sensitive to host hiccups, see the median",
           "reference_wiring_ms_per_step": ref_ms,
           "layout": "channel-last rows (train_rows.py): BatchNorm+ReLU(+max-pool) forward/backward by csrc/train_rows.cu, "
                     "GEMMs by cuBLAS with TF32 allowed, indices by libpn2_b200; the next batch's sampling chain is "
                     "prefetched under this step's backward",
           "grad_bytes": tr.bucket.flat.numel() * 4, "allreduce_us_alone": ar_us,
           "allreduce": "3 chunks launched from gradient hooks during backward (sharding.FlatGradAllReduce.enable_overlap)"
                        if world > 1 else None,
           "parity": par}
    if par is not None:
        par["all_ranks_ok"] = par["ok"]
    del tr, net, pc, pcs
    torch.cuda.empty_cache()
    return [rec]


def sub_config(args, which, rank, local_rank, world, K):
    """A short run of another BASELINE.json configuration: resident throughput + the parity gate."""
    import torch
    if which == 4:
        return train_config(args, rank, local_rank, world, K)
    recs = []
    sizes = (100000, 150000, 200000) if which == 5 else (None,)
    for n in sizes:
        cfg = config_of(args, which)
        if n is not None:
            cfg["points"] = n
            cfg["what"] = cfg["what"].replace("200 000", "{:,}".format(n).replace(",", " "))
        r = Runner(args, cfg, rank, local_rank, world).build()
        dt, launches, _ = r.resident(K, 3)
        par = None if args.no_parity else r.gate()
        recs.append({"config": which, "workload": cfg["what"], "scenes_per_gpu": cfg["batch"], "points": cfg["points"],
                     "precision": cfg["precision"], "batches_in_flight": cfg["lanes"], "steps": K,
                     "sampling_mode": getattr(r, "sampling_mode", "latency"),
                     "value": world * cfg["batch"] * K / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / K,
                     "gpu_launches_per_step": launches / K, "parity": par})
        del r
        torch.cuda.empty_cache()
    return recs


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from situation3d_b200 import fused

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    prev_affinity, numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else (None, 0)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.config == 4:
        rec = train_config(args, rank, local_rank, world, args.steps)[0]
        if rank == 0:
            line = {"metric": "Pointnet2Backbone training step scenes/sec (40k pts)", "value": rec["value"], "unit": UNIT,
                    "n_gpus": world, "steps": args.steps, "warmup": 3, "ms_per_step": rec["ms_per_step"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (tf32 GEMMs)",
                    "data": "synthetic", "config": {"workload": rec["workload"], "config_id": 4}, "train": rec}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    cfg = config_of(args)
    if cfg["precision"] == "bf16" and "bf16" not in fused.SA_FORWARD:
        cfg["precision"] = "fp32"
    precision = cfg["precision"]
    pk = peaks()
    B, K, W = cfg["batch"], args.steps, max(args.warmup, 3)

    run = Runner(args, cfg, rank, local_rank, world).build()
    net = run.backbone
    in_bytes = run.host.numel() * 4
    # nvidia-smi is started well before the timed region: its start-up takes driver locks that stall kernel
    # launches for tens of milliseconds
    sampler = ClockSampler(local_rank) if rank == 0 else None
    part = None
    if args.fps_sms > 0:
        from situation3d_b200.streams import SmPartition
        part = SmPartition(args.fps_sms, dev)
        net.sm_partition = part
        net._side_streams.clear()
    run.barrier()
    time.sleep(1.0)                # every rank: let rank 0's nvidia-smi finish starting up
    run.barrier()
    t_begin = time.time()
    dt, launches, host_ms_per_step = run.resident(K, W, part)
    parity = None if args.no_parity else run.gate(scenes=min(2, B))

    e2e_rec, e2e_f32 = None, None
    if not args.no_e2e and not args.no_graphs:
        run.graphs = None                                      # release the resident-run graphs' pools
        compact = precision == "bf16"
        dt_e, h2d, d2h, res0 = run.e2e(K, W, compact)
        api = "situation3d_b200.graphs.BackbonePipeline.submit(%s) -> pinned host results"
        e2e_rec = {"value": world * B * K / dt_e, "unit": UNIT, "ms_per_step": 1e3 * dt_e / K, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h,
                   "input_format": "fp32 xyz + bf16 feature rows (Pointnet2Backbone.pack_point_clouds; results bit-identical "
                                   "to fp32 point_clouds: tests/test_fused_gpu.py::test_compact_input_is_bit_identical)"
                                   if compact else "fp32 point_clouds",
                   "api": api % ("pinned host xyz f32 + feature rows bf16" if compact else "pinned host point_clouds f32"),
                   "host_numa_binding": "rank pinned to its GPU's %d local CPUs before allocating pinned buffers" % numa_cpus
                                        if numa_cpus else None}
        if not args.no_parity:
            # the end-to-end loop's own result (pinned host tensors of the last batch on lane 0) through the same gate
            full = dict(run.out0)
            full.update({k: v for k, v in res0.items()})
            g = parity_gate(full, run.host, run.sd_cpu, cfg, precision, scenes=1)
            e2e_rec["parity_ok"] = bool(g["ok"])
            if parity is not None:
                parity["e2e_result_ok"] = bool(run.reduce(1.0 if g["ok"] else 0.0, "min") > 0.5)
        if compact:
            dt_f, h2d_f, d2h_f, _ = run.e2e(K, W, False)
            e2e_f32 = {"value": world * B * K / dt_f, "unit": UNIT, "ms_per_step": 1e3 * dt_f / K,
                       "h2d_bytes_per_step": h2d_f, "d2h_bytes_per_step": d2h_f, "input_format": "fp32 point_clouds",
                       "api": api % "pinned host point_clouds f32"}
    clocks = sampler.stop(t_begin, time.time()) if sampler else None
    run.barrier()

    rows, ref_cuda, cpu_base = None, None, None
    if world > 1:
        # the per-kernel breakdown, the CPU baseline and the reference-CUDA arm are single-GPU measurements
        args.no_kernel_breakdown = args.no_reference_cuda = args.no_cpu_baseline = True
    with torch.no_grad():
        if rank == 0:
            if not args.no_kernel_breakdown:
                rows = kernel_breakdown(net, run.pool[0], precision, pk, sampling_mode=getattr(run, "sampling_mode", "latency"))
            try:
                ref_cuda = None if args.no_reference_cuda else reference_cuda_arm(run.pool[0], run.sd_cpu, run.out0,
                                                                                  lanes=cfg["lanes"])
            except Exception as ex:   # test infrastructure must not take the bench down
                ref_cuda = {"unavailable": repr(ex)[:200]}
            if not args.no_cpu_baseline:
                if prev_affinity:
                    os.sched_setaffinity(0, prev_affinity)      # the CPU baseline gets every host core again
                from oracle import pn2_oracle as orc
                cores = os.cpu_count() or 1
                torch.set_num_threads(cores)
                sample = run.host[: min(B, 4)].contiguous()
                orc.backbone(sample[:1].contiguous(), run.sd_cpu, oracle_layers(cfg))
                t0 = time.perf_counter()
                orc.backbone(sample, run.sd_cpu, oracle_layers(cfg))
                t1 = time.perf_counter() - t0
                cpu_base = {"value": sample.shape[0] / t1, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "%d of the step's %d scenes, one pass after a 1-scene warm-up; oracle port "
                                      "(PyTorch CPU wiring over oracle/pn2_oracle.c, OpenMP + %d torch threads)"
                                      % (sample.shape[0], B, cores)}
    run.barrier()
    lanes_n = len(run.lanes)
    sampling_mode_used = getattr(run, "sampling_mode", "latency")
    del run
    torch.cuda.empty_cache()

    # ---- the other configurations, short ----
    subs = {}
    if not args.no_sub_configs and args.config == 2:
        for which in (1, 3, 4, 5):
            try:
                subs[str(which)] = sub_config(args, which, rank, local_rank, world,
                                               (max(4, min(K, 12)) if which == 1 else 24) if which in (1, 4)
                                               else 2 * CONFIGS[which]["lanes"])
            except Exception as ex:
                subs[str(which)] = [{"config": which, "error": repr(ex)[:300]}]

    ok = True
    if rank == 0:
        line = {"metric": METRIC, "value": world * B * K / dt, "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": cfg["what"], "config_id": args.config,
                           "scenes_per_gpu": B, "points": cfg["points"], "feature_channels": 129,
                           "precision": precision, "sharding": "by scene, no collective",
                           "batches_in_flight": lanes_n, "cuda_graphs": not args.no_graphs,
                           "sampling_mode": sampling_mode_used,
                           "sm_partition": {"fps": part.sms[0], "main": part.sms[1]} if part else None,
                           "l2": "each input batch is %.0f MB (> 126 MB L2); two batches alternate" % (in_bytes / 1e6)},
                "gpu_launches": launches * world, "gpu_launches_per_step": launches / K,
                "host_enqueue_ms_per_step": host_ms_per_step, "clocks": clocks, "peaks": pk}
        if e2e_rec:
            line["e2e"] = e2e_rec
        if e2e_f32:
            line["e2e_f32_input"] = e2e_f32
        if parity is not None:
            line["parity"] = parity
            ok &= parity["ok"] and parity["all_ranks_ok"] and parity.get("e2e_result_ok", True)
        if rows:
            dom = max(rows, key=lambda r: r["share"])
            traffic = ncu_traffic()
            line["roofline"] = {"kernel": dom["kernel"], "bound": dom["bound"], "achieved": dom["achieved"],
                                "peak": dom["peak"], "unit": dom["unit"], "frac": dom["frac"],
                                "traffic": traffic.get(dom["kernel"] + (":bucketed" if dom.get("variant", "").startswith("bucketed") else "")),
                                "share_of_step": dom["share"], "us_per_launch": dom["us"],
                                "dominance": "largest share of SM-time of a step (launch duration x fraction of the SMs the "
                                             "launch can occupy: the sampling launches run 1-8 CTAs per scene, everything else is "
                                             "counted as the whole GPU); share_serial in roofline_kernels is plain duration",
                                "peak_source": pk["source"] + (" burst" if dom["bound"] == "tensor" else "")}
            for extra in ("rounds_per_s", "sm_share", "variant"):
                if extra in dom:
                    line["roofline"][extra] = dom[extra]
            if "rounds_per_s" in dom:
                line["roofline"]["note"] = ("a chain of dependent argmax rounds: neither HBM nor the tensor pipe bounds it; "
                                            "rounds_per_s is the figure of merit (DESIGN.md 4.1)")
            # whole step against the ceilings of SURVEY.md 8d (12.38 GFLOP and 33.2 MB per 40k-point scene)
            if cfg["points"] == 40000 and cfg["npoints"][0] == 2048:
                sps = world * B * K / dt / world
                line["roofline_step"] = {"tensor_tflops": 12.382634e-3 * sps, "tensor_frac_sustained": 12.382634e-3 * sps / pk["tensor"],
                                         "hbm_gbs": 33.2e-3 * sps, "hbm_frac": 33.2e-3 * sps / pk["hbm"],
                                         "note": "algorithmic FLOPs / bytes per scene x scenes/s per GPU"}
            line["roofline_kernels"] = [
                {k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items()
                 if k in ("kernel", "variant", "bound", "us", "share", "share_serial", "achieved", "unit", "frac", "hbm_gbs",
                          "rounds_per_s", "sm_share", "layer_tflops")}
                for r in rows]
            for r in line["roofline_kernels"]:
                key = r["kernel"] + (":bucketed" if r.get("variant", "").startswith("bucketed") else "")
                if key in traffic:
                    r["traffic"] = traffic[key]
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        if ref_cuda:
            if "value" in ref_cuda:
                ref_cuda["ours_over_reference_cuda"] = {"resident": line["value"] / ref_cuda["value"],
                                                        "e2e": (line["e2e"]["value"] / ref_cuda["value"]) if "e2e" in line else None}
            line["reference_cuda"] = ref_cuda
        if subs:
            line["configs"] = subs
            for recs in subs.values():
                for r in recs:
                    if r.get("parity") is not None:
                        ok &= r["parity"]["ok"] and r["parity"]["all_ranks_ok"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and not ok:
        print("bench: PARITY GATE FAILED (see \"parity\" in the line above)", file=sys.stderr)
        sys.exit(3)


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
