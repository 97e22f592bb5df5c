#!/usr/bin/env python
"""Headline benchmark: pseudo-labelled images/sec of the U-Net ensemble + Inconsistency-Mask
hot path (BASELINE.json metric) on synthetic images.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Workloads (`--config`; seeded random weights, synthetic uint8 images):

    isic5  (default) BASELINE configs[4], the sweep the metric's "1/2/4/8 B200" is quoted on:
           256x256x3, 5 x U-Net alpha=0.5 (171,561 params), sigmoid, strict > 0.5, IM + blanking
    isic2  configs[0]: the same with 2 models          hela   configs[1]: 256x256x1, 2 x alpha=1, 3 sigmoid heads >= 0.5
    suim   configs[2]: 256x256x3, 2 x alpha=2, softmax K=9, argmax     city / city2  configs[3]: 208x416x3, K=35, alpha 1 / 2

One step = one pass of the hot path over `--images-per-step` images PER GPU (weak scaling: the
pool shards by image, no data-path collective; one int64[3] all-reduce of the coverage statistics
per step).  With no `--config` the default workload is measured in full and, at N = 1, every
other config is measured briefly and reported under "configs" of the same JSON line.

Printed JSON line (rank 0): value = whole-job images/s with the inputs resident in HBM; e2e = the
same through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region);
roofline = the dominant kernel by device time measured live with CUDA events on the launching
stream (imk_profile_begin/end); roofline_step = the whole step against the layer-wise HBM bound
of SURVEY.md 8d; roofline_im = the stand-alone fused IM kernel on materialised probabilities (the
>= 70 % of HBM target of BASELINE.md); cpu_baseline = the oracle port of the reference's CPU path
on this host's cores (batch-1 as the reference runs it, plus a batch-64 variant).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pseudo-labelled images/sec (U-Net ensemble+IM)"

CONFIGS = {
    "isic2": dict(H=256, W=256, c=3, K=1, alpha=0.5, act="sigmoid", M=2, kind="binary", strict=1, seed=1,
                  workload="ISIC 256x256x3 uint8, 2 x U-Net alpha=0.5 (171,561 params), sigmoid > 0.5, IM + blanking"),
    "hela": dict(H=256, W=256, c=1, K=3, alpha=1.0, act="sigmoid", M=2, kind="binary", strict=0, seed=2,
                 workload="HeLa 256x256x1 uint8, 2 x U-Net alpha=1 (681,683 params), 3 sigmoid heads >= 0.5, combined IM + blanking"),
    "suim": dict(H=256, W=256, c=3, K=9, alpha=2.0, act="softmax", M=2, kind="multiclass", strict=0, seed=3,
                 workload="SUIM 256x256x3 uint8, 2 x U-Net alpha=2 (2,717,865 params), softmax K=9 argmax, IM + blanking"),
    "city": dict(H=208, W=416, c=3, K=35, alpha=1.0, act="softmax", M=2, kind="multiclass", strict=0, seed=4,
                 workload="Cityscapes 208x416x3 uint8, 2 x U-Net alpha=1 (682,259 params), softmax K=35 argmax, IM + blanking"),
    "city2": dict(H=208, W=416, c=3, K=35, alpha=2.0, act="softmax", M=2, kind="multiclass", strict=0, seed=4,
                  workload="Cityscapes 208x416x3 uint8, 2 x U-Net alpha=2 (2,718,723 params), softmax K=35 argmax, IM + blanking"),
    "isic5": dict(H=256, W=256, c=3, K=1, alpha=0.5, act="sigmoid", M=5, kind="binary", strict=1, seed=5,
                  workload="ISIC scaling sweep 256x256x3 uint8, 5 x U-Net alpha=0.5 (171,561 params), sigmoid > 0.5, IM + blanking"),
}
DEFAULT_CONFIG = "isic5"       # BASELINE.json configs[4]: the sweep the metric's 1/2/4/8-GPU curve is quoted on


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def make_weights(cfg):
    from inconsistencymasks_b200 import weights as W            # pure NumPy: does not load libimk.so
    return [W.init_weights(cfg["c"], cfg["K"], cfg["alpha"], seed=cfg["seed"] + j) for j in range(cfg["M"])]


def planes_of(cfg):
    return 1 if cfg["kind"] == "multiclass" else cfg["K"]


# ------------------------------------------------------------------------------ algorithmic work
def layer_work(cfg, n_images):
    """Algorithmic bytes / flops per U-Net layer for n images: fp16 activations at the reference's own channel
    counts (SURVEY.md 8d 'layerwise bytes'; no padding), uint8 image in, weights counted once."""
    from inconsistencymasks_b200 import weights as W
    plan = [it for it in W.layer_plan(cfg["c"], cfg["K"], cfg["alpha"]) if it[0] == "conv"]
    # resolution level of each of the 24 convs, creation order (unet.py:49-63)
    levels = [0] + [0, 0, 1, 1, 2, 2, 3, 3] + [4, 4] + [3, 3, 3, 2, 2, 2, 1, 1, 1, 0, 0, 0] + [0]
    out = []
    for i, ((_, ks, cin, cout), lvl) in enumerate(zip(plan, levels)):
        px = (cfg["H"] >> lvl) * (cfg["W"] >> lvl) * n_images
        first, last = i == 0, i == len(plan) - 1
        in_b = px * (cin * 1 if first else cin * 2)
        if 11 <= i <= 22 and (i - 11) % 3 == 0:       # decoder entry conv also reads the low-res map (upsample + add)
            in_b += px // 4 * cin * 2
        out_b = px * (cout * 4 if last else cout * 2)
        out.append(dict(layer=i, ks=ks, cin=cin, cout=cout, level=lvl, bytes=in_b + out_b + ks * ks * cin * cout * 2,
                        in_bytes=in_b, out_bytes=out_b, w_bytes=ks * ks * cin * cout * 2, flops=2.0 * px * ks * ks * cin * cout, px=px))
    return out


def kernel_work(work, cfg, kernel, tag):
    """Algorithmic work of one launch.  A block-fused kernel reads its first layer's input and writes its last layer's
    output (the maps in between never leave the SM); `block_head` also holds the output layer and writes one decision
    byte per pixel (fused ensemble path)."""
    if kernel.startswith("block_"):
        n = {"block_front": 3, "block_enc": 2, "block_dec": 3, "block_head": 4}[kernel]
        layers = [work[tag + j] for j in range(n)]
        out_b = layers[-1]["px"] if kernel == "block_head" else layers[-1]["out_bytes"]
        return dict(bytes=layers[0]["in_bytes"] + out_b + sum(l["w_bytes"] for l in layers), flops=sum(l["flops"] for l in layers))
    return work[tag]


def im_bytes_per_image(cfg):
    """px * (4*K*M + c_in + c_out + n_label + 1)."""
    return cfg["H"] * cfg["W"] * (4 * cfg["K"] * cfg["M"] + 2 * cfg["c"] + planes_of(cfg) + 1)


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------ CPU baseline (oracle port)
def _cpu_im(cfg, images, probs, i):
    """The reference's NumPy IM arithmetic + blanking on image i (functions.py:2854-2874 / 2938-2974 / 3029-3061)."""
    from oracle import ref_im
    H, W = cfg["H"], cfg["W"]
    if cfg["kind"] == "multiclass":
        lab, im, im_size, _ = ref_im.im_prediction_multiclass(probs)
        ref_im.blank_multiclass(images[i], lab, im)
    elif cfg["K"] == 3:
        alive, dead, pos, cim, im_size = ref_im.im_prediction_hela(probs)
        ref_im.blank_hela(images[i, ..., 0], alive, dead, np.zeros((H, W, 3), np.uint8), cim)
    else:
        lab, im, im_size, _ = ref_im.im_prediction_binary(probs, 0.5)
        ref_im.blank_binary(images[i], lab, im)
    return im_size


def cpu_reference_pass(cfg, images, weights, threads, batch=1):
    """The per-directory loop of create_pseudo_labels_im_* (functions.py:2844-2887, 2932-2980, 3020-3066) minus PNG
    I/O and circle drawing.  batch == 1 executes it the way the reference does: per image, per model, a batch-1 forward
    (oracle: torch fp32 CPU), then the reference's NumPy IM arithmetic and blanking.  batch > 1 is the variant
    BASELINE.md section 4 asks for: the forward runs on `batch` images at a time, the IM stays per image."""
    import torch
    from oracle import ref_unet
    torch.set_num_threads(threads)
    sizes = []
    for i0 in range(0, images.shape[0], batch):
        chunk = images[i0:i0 + batch]
        probs = [ref_unet.forward(chunk, w, cfg["act"]) for w in weights]
        for j in range(chunk.shape[0]):
            sizes.append(_cpu_im(cfg, images, [p[j] for p in probs], i0 + j))
    return sizes


def time_cpu(cfg, weights, n_images, threads, batch=1, seed=1234):
    rng = np.random.default_rng(seed)
    images = rng.integers(0, 256, size=(n_images, cfg["H"], cfg["W"], cfg["c"]), dtype=np.uint8)
    cpu_reference_pass(cfg, images[:min(batch, 2)], weights, threads, batch)            # warm-up
    t0 = time.perf_counter()
    cpu_reference_pass(cfg, images, weights, threads, batch)
    dt = time.perf_counter() - t0
    return n_images / dt, dt


def config_block(name, cfg, world):
    """`config` of the JSON line: identical on both arms (the driver compares them)."""
    return dict(workload=cfg["workload"], name=name, parallelism=f"image-sharded x{world}",
                l2="inputs and activations per step exceed the 126 MB L2 (no flush needed)")


def run_reference(args, name, cfg, rank, world):
    """`--impl reference`: the oracle port of the reference's CPU path on this box's host cores.  Imports nothing that
    loads libimk.so (inconsistencymasks_b200.weights is pure NumPy)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    weights = make_weights(cfg)
    per_step = args.ref_images_per_step
    rng = np.random.default_rng(99)
    images = rng.integers(0, 256, size=(per_step, cfg["H"], cfg["W"], cfg["c"]), dtype=np.uint8)
    for _ in range(args.warmup):
        cpu_reference_pass(cfg, images[:2], weights, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_pass(cfg, images, weights, threads)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = dict(metric=METRIC, value=value, unit="images/s", impl="reference",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * dt / args.steps,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=config_block(name, cfg, world),
                sample=dict(images_per_step=per_step, note="each step is a bounded sample of the workload (the CPU path is ~1000x slower)"),
                cpu_baseline=dict(value=value, unit="images/s", cores=threads, kind="port",
                                  sample=f"{per_step} images/step x {args.steps} steps, batch-1 torch fp32 forward per model + NumPy IM + blanking"),
                e2e=dict(value=value, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank (and therefore its pinned host buffers, first-touch) to the NUMA node its GPU hangs off: at 8 ranks
    the end-to-end path moves ~18 GB/s per GPU through host memory, and remote-socket buffers cap it."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------ GPU arm
class Workload:
    """Models + device buffers of one config on one GPU; `step()` is one pass of the fused hot path over N images."""

    def __init__(self, cfg, N, dev, rank, engine=None):
        import torch
        from inconsistencymasks_b200 import unet
        self.cfg, self.N, self.dev = cfg, N, dev
        self.weights = make_weights(cfg)
        self.models = [unet.B200UNet(cfg["H"], cfg["W"], cfg["c"], cfg["K"], cfg["alpha"], cfg["act"], w) for w in self.weights]
        if engine:
            for mdl in self.models:
                mdl.set_engine(engine)
        M, H, W, c = cfg["M"], cfg["H"], cfg["W"], cfg["c"]
        self.handles = (C.c_void_p * M)(*[m.handle for m in self.models])
        gen = torch.Generator(device=dev)
        gen.manual_seed(1000 + rank)
        self.images = torch.randint(0, 256, (N, H, W, c), dtype=torch.uint8, device=dev, generator=gen)
        self.img_out = torch.empty_like(self.images)
        self.planes = planes_of(cfg)
        self.labels = torch.empty((self.planes, N, H, W), dtype=torch.uint8, device=dev)
        self.im = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
        self.im_size = torch.empty(N, dtype=torch.int64, device=dev)
        self.pred_size = torch.zeros((self.planes, N), dtype=torch.int64, device=dev)
        self.stream = torch.cuda.current_stream().cuda_stream

    def step(self):
        from inconsistencymasks_b200._lib import lib, check
        cfg = self.cfg
        if cfg["kind"] == "multiclass":
            check(lib.imk_ensemble_im_multiclass(self.handles, cfg["M"], self.images.data_ptr(), self.N, 0, 1, 1, self.img_out.data_ptr(),
                                                 self.labels.data_ptr(), self.im.data_ptr(), self.im_size.data_ptr(), None, self.stream))
        else:
            check(lib.imk_ensemble_im_binary(self.handles, cfg["M"], self.images.data_ptr(), self.N, 0, 0.5, cfg["strict"], 1, 1,
                                             self.img_out.data_ptr(), self.labels.data_ptr(), self.im.data_ptr(), self.im_size.data_ptr(),
                                             self.pred_size.data_ptr(), self.stream))

    def close(self):
        for m in self.models:
            m.close()


class HostWorkload:
    """Pinned host buffers + the host-buffer C-ABI call (imk_pseudo_label_*_host): uploads, kernels and downloads of
    consecutive chunks overlap inside the library; H2D + D2H are inside the timed region."""

    def __init__(self, wl, Ne, chunk, packed=False):
        import torch
        cfg = wl.cfg
        H, W, c = cfg["H"], cfg["W"], cfg["c"]
        self.wl, self.Ne, self.chunk, self.packed = wl, Ne, chunk, packed
        self.h_img = torch.randint(0, 256, (Ne, H, W, c), dtype=torch.uint8).pin_memory()
        self.h_out = torch.empty_like(self.h_img).pin_memory()
        self.h_lab = torch.empty((wl.planes, Ne, H, W), dtype=torch.uint8).pin_memory()
        self.h_im = torch.empty((Ne, H, W), dtype=torch.uint8).pin_memory()
        self.h_sz = torch.empty(Ne, dtype=torch.int64).pin_memory()
        self.h_pred = torch.empty((wl.planes, Ne), dtype=torch.int64).pin_memory()
        self.h2d = Ne * H * W * c
        stats = Ne * 8 * (1 + (0 if cfg["kind"] == "multiclass" else wl.planes))
        self.d2h = Ne * H * W * (c + wl.planes + 1) + stats
        if packed:       # opt-in layout: 0/255 planes as bits, no blanked image (the host holds the image it uploaded)
            self.d2h = Ne * H * W * ((wl.planes if cfg["kind"] == "multiclass" else 0) * 8 + (0 if cfg["kind"] == "multiclass" else wl.planes) + 1) // 8 + stats

    def step(self):
        from inconsistencymasks_b200._lib import lib, check
        wl, cfg = self.wl, self.wl.cfg
        if self.packed and cfg["kind"] == "multiclass":
            check(lib.imk_pseudo_label_multiclass_host_packed(wl.handles, cfg["M"], self.h_img.data_ptr(), self.Ne, 0, 1, 1, None,
                                                              self.h_lab.data_ptr(), self.h_im.data_ptr(), self.h_sz.data_ptr(), None, self.chunk))
        elif self.packed:
            check(lib.imk_pseudo_label_binary_host_packed(wl.handles, cfg["M"], self.h_img.data_ptr(), self.Ne, 0, 0.5, cfg["strict"], 1, 1,
                                                          None, self.h_lab.data_ptr(), self.h_im.data_ptr(), self.h_sz.data_ptr(),
                                                          self.h_pred.data_ptr(), self.chunk))
        elif cfg["kind"] == "multiclass":
            check(lib.imk_pseudo_label_multiclass_host(wl.handles, cfg["M"], self.h_img.data_ptr(), self.Ne, 0, 1, 1, self.h_out.data_ptr(),
                                                       self.h_lab.data_ptr(), self.h_im.data_ptr(), self.h_sz.data_ptr(), None, self.chunk))
        else:
            check(lib.imk_pseudo_label_binary_host(wl.handles, cfg["M"], self.h_img.data_ptr(), self.Ne, 0, 0.5, cfg["strict"], 1, 1,
                                                   self.h_out.data_ptr(), self.h_lab.data_ptr(), self.h_im.data_ptr(), self.h_sz.data_ptr(),
                                                   self.h_pred.data_ptr(), self.chunk))


def kernel_profile(wl, pk, steps=2):
    """Per-kernel device time of `steps` passes (CUDA events on the launching stream) -> rows sorted by time, each with
    its algorithmic GB/s and TFLOP/s; plus the roofline object of the dominant kernel."""
    from inconsistencymasks_b200 import _lib
    from inconsistencymasks_b200._lib import lib
    cfg = wl.cfg
    _lib.profile_begin()
    for _ in range(steps):
        wl.step()
    prof = _lib.profile_end()
    total_ms = sum(p["total_ms"] for p in prof) or 1.0
    chunk = min(int(lib.imk_max_chunk()), wl.N)      # images per kernel launch of the trunk
    work = {w["layer"]: w for w in layer_work(cfg, chunk)}
    rows = []
    for p in sorted(prof, key=lambda p: -p["total_ms"]):
        avg_ms = p["total_ms"] / p["launches"]
        row = dict(kernel=p["name"], layer=p["tag"], launches=p["launches"], avg_us=1e3 * avg_ms, share=p["total_ms"] / total_ms)
        if p["tag"] in work and p["name"].startswith(("conv", "in_conv", "block_")):
            wk = kernel_work(work, cfg, p["name"], p["tag"])
            row.update(gbs=wk["bytes"] / avg_ms / 1e6, tflops=wk["flops"] / avg_ms / 1e9, bytes=wk["bytes"], flops=wk["flops"])
        elif p["name"] in ("ensemble_im", "ensemble_votes"):
            # fused epilogue: reads M fp16 c9 maps (ensemble_im) or M decision bytes (ensemble_votes) + the image,
            # writes image, labels and IM as uint8
            c9_ch = 8 if int(16 * cfg["alpha"]) <= 8 else (int(16 * cfg["alpha"]) + 15) // 16 * 16     # 8-channel maps keep one 16-byte plane
            per_px = (2 * c9_ch if p["name"] == "ensemble_im" else 1) * cfg["M"] + 2 * cfg["c"] + wl.planes + 1
            b = chunk * cfg["H"] * cfg["W"] * per_px
            row.update(gbs=b / avg_ms / 1e6, tflops=0.0, bytes=b, flops=0.0)
        rows.append(row)
    top = rows[0]
    traffic = measured_traffic(f"{top['kernel']}:{top['layer']}", chunk)
    if "gbs" in top:
        ai = top["flops"] / top["bytes"]
        ridge = pk["tf_sustained"] * 1e12 / (pk["hbm"] * 1e9)
        if ai > ridge:
            roof = dict(bound="tensor", achieved=top["tflops"], peak=pk["tf_sustained"], unit="TFLOP/s",
                        frac=top["tflops"] / pk["tf_sustained"], traffic=traffic)
        else:
            roof = dict(bound="hbm", achieved=top["gbs"], peak=pk["hbm"], unit="GB/s", frac=top["gbs"] / pk["hbm"], traffic=traffic)
        roof.update(kernel=top["kernel"], layer=top["layer"], share_of_step=top["share"], avg_us=top["avg_us"],
                    tflops=top["tflops"], gbs=top["gbs"], peak_source=pk["source"] + " (sustained bf16 / copy)",
                    algorithmic_bytes=top["bytes"], images_per_launch=chunk)
    else:
        roof = dict(bound="hbm", achieved=None, peak=pk["hbm"], unit="GB/s", frac=None, traffic=traffic, kernel=top["kernel"],
                    layer=top["layer"], share_of_step=top["share"], avg_us=top["avg_us"], peak_source=pk["source"])
    return rows, roof


def measured_traffic(key, images_per_launch):
    """DRAM bytes of one launch from the committed `ncu --set full` capture -- only when that capture was taken from the
    sources this library was built from (the digest is stored next to the numbers); otherwise null, never a stale figure."""
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(tpath):
        return None
    try:
        from inconsistencymasks_b200 import build as _build
        t = json.load(open(tpath))
        if t.get("sources_digest") != _build.sources_digest():
            return None
        e = t.get("kernels", {}).get(key)
        return float(e["bytes_per_image"]) * images_per_launch if e else None
    except Exception:
        return None


def resident_step(wl, world, stats):
    """One step of the resident-input measurement: the fused ensemble call + the coverage statistic (row a10)."""
    import torch.distributed as dist
    wl.step()
    if stats is not None:
        stats[0] = wl.im_size.sum(); stats[1] = wl.pred_size.sum(); stats[2] = wl.N
        if world > 1:
            dist.all_reduce(stats)          # the one collective: coverage statistics


def time_resident(wl, steps, warmup, world, barrier, stats=None):
    """K timed steps between two device events, barrier + synchronize on both sides; returns ms for all K steps.
    Warm-up steps run the SAME step (including the statistic's small reductions: their first launch loads a module)."""
    import torch
    for _ in range(warmup):
        resident_step(wl, world, stats)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        resident_step(wl, world, stats)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def time_im_kernel(wl, pk, Ni):
    """Stand-alone IM kernel on materialised fp32 probabilities (BASELINE.md target: >= 70 % of HBM); L2 flushed."""
    import torch
    from inconsistencymasks_b200._lib import lib, check
    cfg, dev = wl.cfg, wl.dev
    M, K, H, W, c = cfg["M"], cfg["K"], cfg["H"], cfg["W"], cfg["c"]
    Ni = min(Ni, wl.N)
    probs = [torch.rand((Ni, H, W, K), dtype=torch.float32, device=dev) for _ in range(M)]
    ptrs = (C.c_void_p * M)(*[p.data_ptr() for p in probs])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    im_ms = []
    for it in range(8):
        flush.fill_(it)                         # write a buffer larger than L2 between timed launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if cfg["kind"] == "multiclass":
            check(lib.imk_im_multiclass(ptrs, M, Ni, H, W, K, wl.images.data_ptr(), c, 1, 1, wl.img_out.data_ptr(), wl.labels.data_ptr(),
                                        wl.im.data_ptr(), wl.im_size.data_ptr(), None, wl.stream))
        else:
            check(lib.imk_im_binary(ptrs, M, Ni, H, W, K, 0.5, cfg["strict"], wl.images.data_ptr(), c, 1, 1, wl.img_out.data_ptr(),
                                    wl.labels.data_ptr(), wl.im.data_ptr(), wl.im_size.data_ptr(), wl.pred_size.data_ptr(), wl.stream))
        b.record()
        torch.cuda.synchronize()
        if it >= 3:
            im_ms.append(a.elapsed_time(b))
    im_t = float(np.mean(im_ms))
    bpi = im_bytes_per_image(cfg)
    gbs = Ni * bpi / (im_t * 1e-3) / 1e9
    return dict(kernel="im_multiclass_tma" if cfg["kind"] == "multiclass" else f"im_binary_vec<{K}>", bound="hbm", achieved=gbs,
                peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"], traffic=None, images=Ni, ms=im_t, bytes_per_image=bpi,
                peak_source=pk["source"], note="timed with the stat memsets on the same stream; L2 flushed between launches")


def step_roofline(cfg, value_per_gpu, pk):
    """Whole step against the layer-wise HBM bound of SURVEY.md 8d: every layer of every model round-trips HBM once at the
    reference's channel counts, plus the fused epilogue's image / label / IM bytes.  Block fusion can exceed 1.0."""
    per_model = sum(w["bytes"] for w in layer_work(cfg, 1)[:-1])
    per_image = cfg["M"] * per_model + cfg["H"] * cfg["W"] * (2 * cfg["c"] + planes_of(cfg) + 1)
    gbs = per_image * value_per_gpu / 1e9
    return dict(bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"], layerwise_bytes_per_image=per_image,
                images_per_s_at_peak=pk["hbm"] * 1e9 / per_image, peak_source=pk["source"])


def run_b200(args, name, cfg, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from inconsistencymasks_b200 import _lib, pool

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator comes up: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    N = args.images_per_step
    wl = Workload(cfg, N, dev, rank, args.engine)
    stats = torch.zeros(3, dtype=torch.int64, device=dev)
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        resident_step(wl, world, stats)     # the step that is timed below, statistic included
    # settle: a fresh process on a fresh box sometimes runs its first few steps ~20 % slow (clock ramp, first-touch of the
    # workspaces); keep warming (untimed, at most 8 more steps, counted in `warmup`) until two consecutive steps agree to 2 %
    torch.cuda.synchronize()
    prev = None
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); resident_step(wl, world, stats); e1.record()
        torch.cuda.synchronize()
        cur = allmax(e0.elapsed_time(e1))
        warmup += 1
        if prev is not None and abs(cur - prev) <= 0.02 * prev:
            break
        prev = cur
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ms = allmax(time_resident(wl, args.steps, 0, world, barrier, stats))
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_im, total_pred, total_n = [int(v) for v in stats.tolist()]
    value = world * N * args.steps / (ms / 1e3)

    # ---- e2e: host buffers through the C ABI (pinned memory; H2D + D2H inside the timed region)
    hw = HostWorkload(wl, args.e2e_images, args.e2e_chunk)
    for _ in range(3):
        hw.step()
    barrier()
    e2e_steps = max(2, args.steps)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hw.step()
    torch.cuda.synchronize()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e_value = world * hw.Ne * e2e_steps / e2e_s
    h2d, d2h = hw.h2d, hw.d2h
    del hw
    # the same through the opt-in packed result layout (bit planes, blanking applied by the host to its own image)
    hwp = HostWorkload(wl, args.e2e_images, args.e2e_chunk, packed=True)
    for _ in range(2):
        hwp.step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hwp.step()
    torch.cuda.synchronize()
    e2e_packed = dict(value=world * hwp.Ne * e2e_steps / allmax(time.perf_counter() - t0), unit="images/s", h2d_bytes_per_step=hwp.h2d,
                      d2h_bytes_per_step=hwp.d2h, api="imk_pseudo_label_*_host_packed (bit planes; img_out = NULL)")
    del hwp

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    rows, roof = kernel_profile(wl, pk)
    roof_im = time_im_kernel(wl, pk, args.im_images)

    # ---- CPU baseline: oracle port on a bounded sample (batch-1 as the reference runs it, and the batch-64 variant)
    threads = os.cpu_count() or 1
    cpu = None
    if not args.no_cpu_baseline:
        v, dt = time_cpu(cfg, wl.weights, args.cpu_images, threads)
        cpu = dict(value=v, unit="images/s", cores=threads, kind="port",
                   sample=f"{args.cpu_images} synthetic images in {dt:.1f} s, batch-1 torch fp32 forward per model + NumPy IM + blanking")
        nb = max(64, args.cpu_images // 64 * 64)
        v64, dt64 = time_cpu(cfg, wl.weights, nb, threads, batch=64)
        cpu["batch64"] = dict(value=v64, unit="images/s", cores=threads,
                              sample=f"{nb} synthetic images in {dt64:.1f} s, batch-64 torch fp32 forward per model + NumPy IM per image")

    # ---- the other BASELINE configs, briefly (N = 1 only; the default run stays within minutes)
    others = []
    if world == 1 and args.config is None and not args.no_other_configs:
        wl.close()
        del wl
        torch.cuda.empty_cache()
        for oname, ocfg in CONFIGS.items():
            if oname == name:
                continue
            try:
                others.append(measure_brief(oname, ocfg, dev, pk, args))
            except Exception as e:                      # a config that fails must not take the headline line with it
                others.append(dict(name=oname, error=str(e)[:200]))

    line = dict(metric=METRIC, value=value, unit="images/s", n_gpus=world,
                steps=args.steps, warmup=warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f16", data="synthetic",
                config=config_block(name, cfg, world),
                run=dict(images_per_step_per_gpu=N, engine=args.engine or "default", mean_im_size=pool.mean_im_size(total_im, total_n),
                         numa_node_rank0=numa_node),
                e2e=dict(value=e2e_value, unit="images/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         images_per_step=args.e2e_images, steps=e2e_steps,
                         api="imk_pseudo_label_%s_host (pinned host buffers)" % ("multiclass" if cfg["kind"] == "multiclass" else "binary")),
                e2e_packed=e2e_packed, gpu_launches=int(launches), clocks=clocks, roofline=roof, roofline_step=step_roofline(cfg, value / world, pk),
                roofline_im=roof_im, cpu_baseline=cpu,
                kernels=[{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ("bytes", "flops")}
                         for r in rows[:12]],
                configs=others)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_brief(name, cfg, dev, pk, args):
    """One BASELINE config at reduced step count: resident value, e2e, roofline of the dominant kernel, stand-alone IM."""
    import torch
    big = cfg["alpha"] >= 2.0 or cfg["K"] >= 35
    N = 1024 if big else 2048
    wl = Workload(cfg, N, dev, 0, args.engine)
    noop = lambda: torch.cuda.synchronize()
    ms = time_resident(wl, 3, 3, 1, noop)
    value = N * 3 / (ms / 1e3)
    hw = HostWorkload(wl, N, args.e2e_chunk)
    for _ in range(2):
        hw.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        hw.step()
    torch.cuda.synchronize()
    e2e = N * 2 / (time.perf_counter() - t0)
    del hw
    rows, roof = kernel_profile(wl, pk, steps=1)
    roof_im = time_im_kernel(wl, pk, 256 if big else 512)
    out = dict(name=name, workload=cfg["workload"], value=value, e2e=e2e, unit="images/s", images_per_step=N,
               roofline={k: roof.get(k) for k in ("kernel", "layer", "bound", "achieved", "peak", "unit", "frac", "share_of_step", "avg_us")},
               roofline_step=step_roofline(cfg, value, pk)["frac"], roofline_im=roof_im["frac"],
               kernels=[(r["kernel"], r["layer"], round(r["share"], 3)) for r in rows[:5]])
    wl.close()
    del wl
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS))
    ap.add_argument("--images-per-step", type=int, default=None)
    ap.add_argument("--e2e-images", type=int, default=None)
    ap.add_argument("--e2e-chunk", type=int, default=0, help="images per chunk of the host pipeline; 0 = the library's default schedule")
    ap.add_argument("--im-images", type=int, default=512)
    ap.add_argument("--cpu-images", type=int, default=128)
    ap.add_argument("--ref-images-per-step", type=int, default=32)
    ap.add_argument("--engine", default=None, choices=[None, "direct", "tcgen05", "fused"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    name = args.config or DEFAULT_CONFIG
    cfg = CONFIGS[name]
    big = cfg["alpha"] >= 2.0 or cfg["K"] >= 35
    if args.images_per_step is None:
        args.images_per_step = 1024 if big else 2048
    if args.e2e_images is None:
        args.e2e_images = 2048 if big else 4096
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, name, cfg, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_b200(args, name, cfg, rank, local_rank, world)


if __name__ == "__main__":
    main()
