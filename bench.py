"""bench.py -- headline benchmark of the hot path: Soft-IntroVAE z=1200 training volumes/sec.

    python bench.py --gpus N --steps K --warmup W            # our arm (libsivae.so on B200)
    python bench.py --impl reference --steps K --warmup W    # reference arm: the reference's CPU path

Workload (BASELINE.json configs[2], z-1200main.py:158,190-202): ``SoftIntroVAE(64, [[64,1,2],[128,1,2],[256,2,2]])``,
synthetic volumes 1x80x96x80 in [0,1], local batch 8 per GPU, one "step" = one full E-update + D-update of
utils/my_trainer.py:236-325 (13 forward passes, 2 backward passes, 2 Adam steps), random-init weights
(init_weights_he under seed 77).  N>1: one process per GPU (torchrun), local batch fixed (weak scaling),
bucketed NCCL gradient all-reduce overlapped with backward, replica-local BatchNorm statistics.

One JSON line on stdout (rank 0).  ``value`` = volumes/s with inputs resident in HBM; ``e2e`` = the same
through the public API with pinned-host inputs copied in and the loss scalars read back every step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

BLOCK_SETTING = [[64, 1, 2], [128, 1, 2], [256, 2, 2]]
IN_CH = 64
VOL = (80, 96, 80)
LOCAL_BATCH = 8
GFLOP_PER_VOLUME_STEP = 7270.4       # 32 network passes x 227.2 GFLOP (SURVEY.md section 8d / BASELINE.md section 2)
METRIC = "train volumes/sec (Soft-IntroVAE z=1200, 80x96x80)"
# roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel on its top shape,
# conv3_kd3_kernel 64->64 @ 8x80x96x80, algorithmic 2 x 629.1 MB) is READ from the committed summary of the ncu
# --set full capture of the current kernel (tools/gpu/ncu_profile.sh writes it); absent file -> null.
TOP_KERNEL_PROFILE = os.path.join(ROOT, "profiles", "top_kernel_ncu.json")
HW_FLOP_FACTOR = {"conv3_igemm": 1.0, "conv3_wgrad": 1.0, "upconv3_fprop": 8.0 / 27.0, "upconv3_dgrad": 8.0 / 27.0,
                  "upconv3_wgrad": 8.0 / 27.0}     # executed / reference-equivalent MACs (Upsample folded: 27 -> 8 taps)


def _top_kernel_profile():
    """-> (traffic bytes per launch or None, note)."""
    if not os.path.isfile(TOP_KERNEL_PROFILE):
        return None, "no ncu capture of this build committed (profiles/top_kernel_ncu.json absent)"
    with open(TOP_KERNEL_PROFILE) as f:
        d = json.load(f)
    note = (f"ncu --set full, {d.get('kernel')} {d.get('shape')}, launch {d.get('duration_us')} us, tensor pipe "
            f"{d.get('tensor_pipe_active_pct')} % active, source commit {d.get('commit')}; algorithmic bytes 1.258e9; "
            f"profiles/{d.get('summary')}")
    return d.get("dram_bytes_per_launch"), note


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("bf16_tflops_sustained", 1376.2)), float(d.get("hbm_gbs", 6546.9)), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms.  The sampler is started before the warm-up (nvidia-smi
    needs about a second to deliver its first line) and ``mark()`` is called when the timed region begins: only
    samples from then on are reported (if the region is shorter than one period, the samples taken under the same
    load during the warm-up are used and the result says so)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index, self.first = None, [], index, 0

    def mark(self):
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        window = self.lines[self.first:]
        note = "timed region"
        if len(window) < 2:
            window, note = self.lines, "warm-up + timed region (timed region shorter than two sampling periods)"
        for ln in window:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": note}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU path (torch fp32, oneDNN) via the oracle port
# --------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(vol=VOL, batch=1, threads=None, workload="z1200"):
    """One full Soft-IntroVAE E+D iteration of the headline net (or, workload "fc600", of the FC-latent variant
    mymodel.SoftIntroVAE(32,64,128,256,600)) on the host cores (fp32, gradients for both phases; the Adam update
    itself is negligible).  /root/reference does not exist on the GPU box, so this is the oracle port
    (oracle/sivae_oracle.py, validated against the reference in tests/)."""
    from oracle import sivae_oracle as O
    torch.set_num_threads(threads or os.cpu_count() or 1)
    torch.manual_seed(77)
    import sivae_b200
    if workload == "fc600":
        small = O.FcCfg(*FC600["chans"], FC600["z_ch"], (1, 1, 1))
        cfg = O.FcCfg(*FC600["chans"], FC600["z_ch"], FC600["grid"])
        nets = {}
        for c in (small, cfg):                                     # parameter holders only (never run on CPU)
            net = sivae_b200.mymodel.SoftIntroVAE(*FC600["chans"], FC600["z_ch"], latent_grid=c.grid)
            net.apply(sivae_b200.init_weights_he)
            nets[c.grid] = {k: v.detach().clone() for k, v in net.state_dict().items()}

        def step(v=vol):
            c = cfg if tuple(v) == tuple(vol) else small
            d, h, w = (16 * g for g in c.grid)
            real = torch.rand(batch, 1, d, h, w)
            noise = torch.randn(batch, FC600["z_ch"])
            eps = [torch.randn(batch, FC600["z_ch"]) for _ in range(5)]
            O.soft_intro_step_grads(nets[c.grid], c, real, noise, eps, None, O.StepHyper(scale=8.0 / (80 * 96 * 80)))
        return step
    cfg = O.NetCfg.soft_intro(IN_CH, BLOCK_SETTING)
    net = sivae_b200.SoftIntroVAE(IN_CH, BLOCK_SETTING)          # parameter holder only (never run on CPU)
    net.apply(sivae_b200.init_weights_he)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}

    def step(v=vol):
        d, h, w = v
        real = torch.rand(batch, 1, d, h, w)
        noise = torch.randn(batch, 1, d // 8, h // 8, w // 8)
        eps = [torch.randn(batch, 1, d // 8, h // 8, w // 8) for _ in range(5)]
        O.soft_intro_step_grads(sd, cfg, real, noise, eps, None, O.StepHyper())
    return step


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    fc = getattr(args, "workload", "z1200") == "fc600"
    batch = 2 if fc else 1            # the FC-latent variant normalises over B*150 values at the latent grid: keep B >= 2
    step = cpu_reference_step_factory(threads=threads, workload="fc600" if fc else "z1200", batch=batch)
    for _ in range(args.warmup):
        step((16, 16, 16) if fc else (16, 24, 16))   # warm-up on a reduced volume keeps the run bounded
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.steps * batch / dt
    sample = (f"{args.steps} x ({batch} volume(s) 80x96x80, one full E+D iteration, "
              f"{'mymodel.SoftIntroVAE(32,64,128,256,600)' if fc else 'headline net'}, fp32 torch-CPU/oneDNN, "
              f"{threads} threads); warm-up on a reduced volume")
    line = {"impl": "reference", "metric": ("train volumes/sec (Soft-IntroVAE FC-latent z=600, 80x96x80)" if fc else METRIC),
            "value": val, "unit": "volumes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("600z_main.py mymodel.SoftIntroVAE(32,64,128,256,600) FC-latent variant, 80x96x80, one "
                                    "trainer_fc E+D train step" if fc else
                                    "z-1200main.py Soft-IntroVAE z=1200, 80x96x80, one E+D train step"),
                       "local_batch": batch, "device": "host CPU"},
            "cpu_baseline": {"value": val, "unit": "volumes/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
FC600 = dict(chans=(32, 64, 128, 256), z_ch=600, grid=(5, 6, 5), batch=4)


def fc_gflop_per_volume_step(chans, z_ch, grid):
    """Reference-equivalent FLOPs of one trainer_fc iteration per volume (direct-convolution count, as SURVEY 8d):
    13 encoder-sized and 19 decoder-sized passes (fwd 5E+8D, dgrad 5E+7D, wgrad 3E+4D)."""
    c1, c2, c3, c4 = chans
    s = grid[0] * grid[1] * grid[2]
    v = [s * 8 ** i for i in range(5)]                        # voxels at latent .. full resolution
    conv = lambda ci, co, vox: 2.0 * 27 * ci * co * vox        # noqa: E731
    enc = (conv(1, c1, v[4]) + conv(c1, c1, v[4]) + conv(c1, c1, v[3]) + conv(c1, c2, v[3]) + conv(c2, c2, v[2])
           + conv(c2, c3, v[2]) + 3 * conv(c3, c3, v[1]) + conv(c3, c4, v[0]) + 2 * conv(c4, c4, v[0])
           + 2.0 * c4 * s * 2 * z_ch)
    dec = (2.0 * z_ch * c4 * s + 3 * conv(c4, c4, v[0]) + conv(c4, c3, v[1]) + 3 * conv(c3, c3, v[1])
           + conv(c3, c2, v[2]) + conv(c2, c2, v[2]) + conv(c2, c1, v[3]) + conv(c1, c1, v[3]) + conv(c1, c1, v[4])
           + conv(c1, 1, v[4]))
    return (13 * enc + 19 * dec) / 1e9


def blob_volumes(n, vol, seed=1234):
    """Smooth 'brain-like' synthetic volumes in [0,1] on a zero background (the value range of BrainDataset._preprocess,
    utils/data_load.py:25-30; skull-stripped-MRI look): an ellipsoid with a low-frequency texture."""
    d, h, w = vol
    g = torch.Generator().manual_seed(seed)
    zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, d), torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    out = torch.empty(n, 1, d, h, w)
    for i in range(n):
        c = (torch.rand(3, generator=g) - 0.5) * 0.3
        r = 0.55 + 0.25 * torch.rand(3, generator=g)
        body = (((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2) < 1.0
        f = 2.0 + 4.0 * torch.rand(3, generator=g)
        tex = 0.5 + 0.25 * torch.sin(f[0] * zz * 3.14) * torch.cos(f[1] * yy * 3.14) + 0.2 * torch.sin(f[2] * xx * 3.14)
        out[i, 0] = (tex * body).clamp(0, 1)
    return out


def measure_l_shape(dev, steps=3, vol=(160, 192, 160), batch=2):
    """volumes/s of the headline net on the L-shape (160x192x160, same fully-convolutional net, latent
    20x24x20 = 9600), inputs resident in HBM, whole-step CUDA graph + FusedAdam, CUDA-event timing."""
    import sivae_b200
    from sivae_b200 import trainer as T
    torch.manual_seed(77)
    net = sivae_b200.SoftIntroVAE(IN_CH, BLOCK_SETTING)
    net.apply(T.init_weights_he)
    net.to(dev).train()
    opt_e = sivae_b200.FusedAdam(net.encoder.parameters(), lr=2e-4)
    opt_d = sivae_b200.FusedAdam(net.decoder.parameters(), lr=2e-4)
    d, h, w = vol
    # smooth volumes: from a random init on uniform noise the recipe itself is unstable at this size (latent 9600: the
    # fp32 oracle reaches rec_kl ~ 2e17 at step 3, profiles/r02_lshape_probe.txt), and a NaN loss is not a measurement
    real = blob_volumes(batch, vol).to(dev)
    noise = torch.randn(batch, 1, d // 8, h // 8, w // 8, device=dev)
    g = sivae_b200.graph.GraphedTrainStep(net, opt_e, opt_d, real, noise, T.StepHyper(), warmup=2)
    for _ in range(2):
        g(real, noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = g(real, noise)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    gflop = GFLOP_PER_VOLUME_STEP * (d * h * w) / float(VOL[0] * VOL[1] * VOL[2])
    if not float(out["lossE"]) == float(out["lossE"]):
        raise RuntimeError("NaN loss in the L-shape measurement")
    return {"workload": f"same net, {d}x{h}x{w} ({d * h * w} voxels, latent {d // 8}x{h // 8}x{w // 8}), local batch {batch}",
            "data": "synthetic smooth volumes in [0,1] on a zero background",
            "value": batch / (ms / 1e3), "unit": "volumes/s", "ms_per_step": ms, "steps": steps,
            "gflop_per_volume_step": gflop, "whole_step_tflops": batch / (ms / 1e3) * gflop / 1e3,
            "lossE": float(out["lossE"])}


def run_ours(args):
    import torch.distributed as dist
    import sivae_b200
    from sivae_b200 import kernels as K, trainer as T, parallel as P, functional as F

    rank, world, local_rank = P.init_distributed("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    K.device_check()
    sampler = ClockSampler(local_rank) if rank == 0 and not args.no_clocks else None
    if sampler:
        sampler.start()                                     # early: nvidia-smi takes ~1 s to deliver its first sample
    B = args.batch
    if args.global_batch:
        # BASELINE configs[3] as stated (main_DataParallel.py:609: global batch 64 over 2/4/8 GPUs): strong scaling
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        B = args.global_batch // world
    D, H, W = args.vol
    fc = args.workload == "fc600"
    if fc and tuple(args.vol) != VOL:
        raise SystemExit("the FC-latent variant hard-codes 80x96x80 inputs (mymodel.py:125)")
    torch.manual_seed(77)                                   # identical init on every rank
    if fc:
        net = sivae_b200.mymodel.SoftIntroVAE(*FC600["chans"], FC600["z_ch"])
    else:
        net = sivae_b200.SoftIntroVAE(IN_CH, BLOCK_SETTING)
    net.apply(T.init_weights_he)
    net.to(dev).train()
    # N = 1: the whole step is one CUDA graph.  N > 1: three graphs with the two NCCL gradient all-reduces issued
    # between the replays (parallel.FlatGradReducer; graph.GraphedTrainStep).  --no-graph: eager step, and for
    # N > 1 the bucketed reducer that overlaps NCCL with backward (parallel.GradReducer).
    use_graph = not args.no_graph
    if args.torch_adam:
        opt_e = torch.optim.Adam(net.encoder.parameters(), lr=2e-4, capturable=use_graph)
        opt_d = torch.optim.Adam(net.decoder.parameters(), lr=2e-4, capturable=use_graph)
    else:   # fused multi-tensor Adam + bf16 weight re-pack (one C-ABI call per phase)
        opt_e = sivae_b200.FusedAdam(net.encoder.parameters(), lr=2e-4)
        opt_d = sivae_b200.FusedAdam(net.decoder.parameters(), lr=2e-4)
    Reducer = P.FlatGradReducer if use_graph else P.GradReducer
    red_e = Reducer(net.encoder.parameters()) if world > 1 else None
    red_d = Reducer(net.decoder.parameters()) if world > 1 else None
    hp = T.StepHyper(scale=sivae_b200.trainer_fc.SCALE) if fc else T.StepHyper()
    torch.manual_seed(1234 + rank)                          # disjoint synthetic shards / noise per rank
    F.manual_seed(1234 + rank)
    real_host = torch.rand(B, 1, D, H, W).pin_memory()
    noise_host = (torch.randn(B, FC600["z_ch"]) if fc else torch.randn(B, 1, D // 8, H // 8, W // 8)).pin_memory()
    real_dev, noise_dev = real_host.to(dev), noise_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_eager():
        return T.soft_intro_train_step(net, real_dev, noise_dev, opt_e, opt_d, hp, red_e, red_d)

    graphed = None
    graph_note = None
    if use_graph:
        # whole-step CUDA graph (sivae_b200.graph): ~1500 launches per step collapse into one replay (three replays
        # + two NCCL calls with N > 1).  If the capture is refused the step runs eagerly.
        try:
            graphed = sivae_b200.graph.GraphedTrainStep(net, opt_e, opt_d, real_dev, noise_dev, hp, warmup=2,
                                                        reducer_e=red_e, reducer_d=red_d)
        except Exception as ex:  # noqa: BLE001
            graph_note = f"capture failed, ran eagerly: {type(ex).__name__}: {str(ex)[:120]}"
            graphed = None
            torch.cuda.synchronize()
        if world > 1:
            ok = torch.tensor([1 if graphed is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok) == 0:
                graphed = None

    def step_resident():
        if graphed is not None:
            return graphed(real_dev, noise_dev)
        return step_eager()

    def step_e2e():
        if graphed is not None:
            terms = graphed(real_host, noise_host)                     # pinned host -> captured input buffers
        else:
            real = real_host.to(dev, non_blocking=True)
            noise = noise_host.to(dev, non_blocking=True)
            terms = T.soft_intro_train_step(net, real, noise, opt_e, opt_d, hp, red_e, red_d)
        res = torch.stack([terms["lossE"], terms["lossD"]]).cpu()      # device->host read of the step's result
        return res

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, out

    for _ in range(max(args.warmup, 1)):
        step_resident()
    if sampler:
        sampler.mark()
    n0 = K.launch_count()
    with K.KernelTimer() as kt:
        ms, terms = timed(step_resident, args.steps)
    launches = K.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    lossE, lossD = float(terms["lossE"]), float(terms["lossD"])
    if graphed is not None:
        # a replayed graph cannot carry per-kernel events: time the dominant kernels live in one eager step of
        # the same process on the same inputs (CUDA events on the launch stream), and count its launches
        # ... on ONE stream: with the default two-stream issue a kernel's event bracket would also contain whatever the
        # other stream ran meanwhile, and the roofline figure is about the kernel alone
        n0 = K.launch_count()
        saved_streams = (T.TWO_STREAMS, F.WGRAD_STREAM)
        T.TWO_STREAMS, F.WGRAD_STREAM = False, False
        try:
            with K.KernelTimer() as kt:
                step_eager()
                torch.cuda.synchronize()
        finally:
            T.TWO_STREAMS, F.WGRAD_STREAM = saved_streams
        launches = (K.launch_count() - n0) * args.steps
    ksum = kt.summary()
    if rank == 0 and args.kernel_table:
        # per kernel family / per shape table of the live CUDA-event timings (profiles/*_kernel_table.txt)
        rows = []
        for name, d in ksum.items():
            for shape, v in d["by_shape"].items():
                rows.append((v["ms"], name, shape, v["launches"], v["work"]))
        with open(args.kernel_table, "w") as f:
            f.write("ms_total  launches  ms_each  TFLOP/s(ref-equiv)  kernel  shape(N,D,H,W,Ci,Co)\n")
            for ms_, name, shape, n_, work in sorted(rows, reverse=True):
                f.write(f"{ms_:8.3f}  {n_:3d}  {ms_ / n_:8.4f}  {work / (ms_ * 1e-3) / 1e12:8.1f}  {name}  {shape}\n")
    for _ in range(1):
        step_e2e()
    ms_e2e, res = timed(step_e2e, args.steps)
    if not (lossE == lossE and lossD == lossD):
        raise SystemExit("NaN loss in the benchmark step")
    params_identical = None
    if world > 1:
        # replicas must hold bit-identical parameters after the run (identical init + averaged gradients + the same
        # update): fp64 checksums (sum, sum of squares) of every parameter, max over ranks == min over ranks
        chk = torch.stack([torch.stack([p_.detach().double().sum(), p_.detach().double().pow(2).sum()])
                           for p_ in net.parameters()]).flatten()
        mx, mn = chk.clone(), chk.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        params_identical = bool(torch.equal(mx, mn)) and bool(torch.isfinite(chk).all())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    vols = world * B * args.steps
    value = vols / (ms / 1e3)
    e2e = vols / (ms_e2e / 1e3)
    vol_scale = (D * H * W) / float(VOL[0] * VOL[1] * VOL[2])     # FLOPs scale with the voxel count (fully convolutional)
    gflop_step = (fc_gflop_per_volume_step(FC600["chans"], FC600["z_ch"], FC600["grid"]) if fc
                  else GFLOP_PER_VOLUME_STEP * vol_scale)
    peak_tf, peak_gbs, peak_src = _peaks()
    dom = ksum.get("conv3_igemm", dict(launches=0, ms=0.0, work=0.0, by_shape={}))
    achieved = dom["work"] / (dom["ms"] * 1e-3) / 1e12 if dom["ms"] > 0 else 0.0
    top_shape = max(dom["by_shape"].items(), key=lambda kv: kv[1]["ms"]) if dom["by_shape"] else None
    wg = ksum.get("conv3_wgrad", dict(launches=0, ms=0.0, work=0.0))
    # every tensor-core 3x3x3 conv kernel of the step together, in EXECUTED (hardware) FLOPs: tensor-pipe utilisation
    hw_flop = sum(ksum[k]["work"] * f for k, f in HW_FLOP_FACTOR.items() if k in ksum)
    hw_ms = sum(ksum[k]["ms"] for k in HW_FLOP_FACTOR if k in ksum)
    hw_tflops = hw_flop / (hw_ms * 1e-3) / 1e12 if hw_ms > 0 else 0.0
    traffic, traffic_note = _top_kernel_profile()
    roofline = {"bound": "tensor",
                "kernel": "3x3x3 conv fprop/dgrad on tcgen05 (conv3_kd3_kernel + conv3_igemm_kernel, all layer shapes)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": f"{peak_src} bf16_tflops_sustained",
                "traffic": traffic if tuple(args.vol) == VOL and B == LOCAL_BATCH and not fc else None,
                "traffic_note": traffic_note,
                # flat per-kernel figures (dominant kernel on its top shape; all tensor-core conv kernels together)
                "top_ms": top_shape[1]["ms"] / top_shape[1]["launches"] if top_shape else None,
                "top_tflops": top_shape[1]["work"] / (top_shape[1]["ms"] * 1e-3) / 1e12 if top_shape else None,
                "top_frac": (top_shape[1]["work"] / (top_shape[1]["ms"] * 1e-3) / 1e12 / peak_tf) if top_shape else None,
                "hw_tflops": hw_tflops, "all_tc_conv_frac": hw_tflops / peak_tf, "all_tc_conv_ms": hw_ms,
                "launches": dom["launches"],
                "share_of_step": dom["ms"] / (ms / args.steps if graphed is not None else ms) if ms else None,
                "top_shape": ({"NDHWCiCo": list(top_shape[0]),
                               "tflops": top_shape[1]["work"] / (top_shape[1]["ms"] * 1e-3) / 1e12,
                               "ms_per_launch": top_shape[1]["ms"] / top_shape[1]["launches"]} if top_shape else None),
                "wgrad": {"tflops": wg["work"] / (wg["ms"] * 1e-3) / 1e12 if wg["ms"] else 0.0,
                          "share_of_step": wg["ms"] / (ms / args.steps if graphed is not None else ms) if ms else None,
                          "launches": wg["launches"]},
                "timing": ("CUDA events around every launch of the kernel in one eager step of the same run"
                           if graphed is not None else "CUDA events around every launch inside the timed region"),
                "whole_step_tflops": value / world * gflop_step / 1e3}
    line = {"metric": ("train volumes/sec (Soft-IntroVAE FC-latent z=600, 80x96x80)" if fc else METRIC), "value": value,
            "unit": "volumes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": ("600z_main.py mymodel.SoftIntroVAE(32,64,128,256,600) FC-latent variant, 80x96x80, "
                                    "one trainer_fc E+D train step incl. 2 Adam steps" if fc else
                                    "z-1200main.py Soft-IntroVAE(64,[[64,1,2],[128,1,2],[256,2,2]]) "
                                    f"latent {D // 8}x{H // 8}x{W // 8} (z={D * H * W // 512}), {D}x{H}x{W}, "
                                    "one E+D train step incl. 2 Adam steps"),
                       "local_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": "per-step working set is tens of GB >> 126 MB L2 (inputs larger than L2)",
                       "gflop_per_volume_step": gflop_step, "cuda_graph": graphed is not None,
                       "streams": {"independent_passes_on_two_streams": bool(T.TWO_STREAMS),
                                   "weight_gradients_on_own_stream": bool(F.WGRAD_STREAM)},
                       "cuda_graph_note": graph_note},
            "clocks": clocks, "e2e": {"value": e2e, "unit": "volumes/s", "ms_per_step": ms_e2e / args.steps,
                                      "h2d_bytes_per_step": real_host.numel() * 4 + noise_host.numel() * 4,
                                      "d2h_bytes_per_step": int(res.numel() * 4)},
            "gpu_launches": int(launches), "roofline": roofline, "loss": {"lossE": lossE, "lossD": lossD}}
    if world > 1:
        line["params_identical"] = params_identical
    if args.global_batch:
        line["scaling"] = "strong"
        line["config"]["global_batch_mode"] = f"--global-batch {args.global_batch}: local batch {B} on {world} ranks"
    if world == 1 and not fc and not args.no_lshape and tuple(args.vol) == VOL:
        # second, clearly named entry: the "~5 M voxel" point of BASELINE.json's metric -- the same net on the original
        # 160x192x160 resolution (4,915,200 voxels, latent 20x24x20), batch 2, same whole-step graph + FusedAdam
        del graphed
        net = opt_e = opt_d = None
        torch.cuda.empty_cache()
        try:
            line["l_shape"] = measure_l_shape(dev, steps=min(args.steps, 5))
        except Exception as ex:  # noqa: BLE001
            line["l_shape"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
    if fc:
        line["roofline"]["traffic"] = None
        line["roofline"]["note"] = ("channel counts below 64 run zero-padded to 64 (tcgen05 tile width): the 32-channel "
                                    "full-resolution layers execute 4x their reference-equivalent FLOPs; 'achieved' "
                                    "counts executed (padded) FLOPs of the conv kernels, 'whole_step_tflops' "
                                    "reference-equivalent ones")
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = 2 if fc else 1
        step = cpu_reference_step_factory(threads=threads, workload=args.workload, batch=cb)
        step((16, 16, 16) if fc else (16, 24, 16))
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": cb / dt, "unit": "volumes/s", "cores": threads, "kind": "port",
                                "sample": f"{cb} volume(s) 80x96x80, one full E+D iteration of "
                                          f"{'mymodel.SoftIntroVAE(32,64,128,256,600)' if fc else 'the headline net'}, fp32 "
                                          f"torch-CPU/oneDNN on {threads} threads ({dt:.1f} s)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        graphed = None                      # graphs (and any collectives captured in them) go before the process group
        torch.cuda.synchronize()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=LOCAL_BATCH, help="local batch per GPU (headline: 8)")
    ap.add_argument("--vol", type=int, nargs=3, default=list(VOL), metavar=("D", "H", "W"),
                    help="volume extents (headline 80 96 80; 160 192 160 = the ~5M-voxel L-shape, use --batch 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="no nvidia-smi clock sampler child (runs under ncu)")
    ap.add_argument("--no-lshape", action="store_true", help="skip the second (160x192x160, batch 2) measurement")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="fixed GLOBAL batch split over the ranks (BASELINE configs[3]: 64 -> local 32/16/8 at 2/4/8 "
                         "GPUs, strong scaling); default 0 = fixed local batch (--batch), weak scaling")
    ap.add_argument("--workload", default="z1200", choices=["z1200", "fc600"],
                    help="z1200: the headline conv-latent net (BASELINE config 3); fc600: the FC-latent variant "
                         "mymodel.SoftIntroVAE(32,64,128,256,600) of 600z_main.py (BASELINE config 2, batch 4)")
    ap.add_argument("--kernel-table", default=None, help="write the per-kernel/per-shape timing table here")
    ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam instead of the fused optimiser")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of as one CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
