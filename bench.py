#!/usr/bin/env python
"""Benchmark of the AA-CLIP inference hot path on B200 (BASELINE.json: anomaly maps/s, 336 px, ViT-L/14).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch-per-gpu B] [--impl b200|reference]
                    [--sweep] [--text] [--no-cpu-baseline] [--no-e2e] [--no-gpu-baseline] [--no-drop-in]

A step = one pass of the hot path over one batch of synthetic images: ViT-L/14-336 visual encoder with the 6
residual adapters, 4 taps -> ln_post -> seg/det projections -> L2 normalise -> anchor similarity -> gaussian
blur + bilinear upsample -> level-summed 336x336 anomaly maps and image scores (test.py:80-93), through
aaclip_forward_fused (device-resident inputs: `value`) and aaclip_submit_host / aaclip_wait_host (pinned host buffers,
H2D + D2H inside the timed region: `e2e`).  N > 1: one process per GPU (torchrun), batch sharded data-parallel,
one NCCL all-gather of the image scores per step, max-over-ranks timing.  Per-GPU batch: 64 (BASELINE.json
configs[1]); at N = 8 the default is 128 per GPU = global batch 1024 (configs[2]).

Extra legs on rank 0 at N = 1 (all reported inside the one JSON line):
  roofline / roofline_head   per-kernel-class CUDA events of the same step; the head (aaclip_anomaly_head) timed as a
                             replayed CUDA graph of back-to-back calls, bf16 and fp32 tokens
  drop_in                    the reference's call sequence on the drop-in classes: model(image) + 4 x
                             calculate_similarity_map + cat + sum (test.py:80-93), fp32 and bf16 tokens
  torch_gpu_baseline         the oracle (stock PyTorch ops, what the reference runs) on the same B200: fp32 with TF32
                             off / on, and bf16 autocast - the like-for-like GPU baseline of SURVEY 8(d)
  cpu_baseline               the oracle on the host cores: 4 threads (the reference's pin, test.py:28-35) and all
                             cores, batch 1 and batch 8
  --sweep                    BASELINE.json configs[4]: batch 1 .. 2048 with the roofline fractions
  --text                     BASELINE.json configs[3]: text-anchor path + similarity against cached patch features

--impl reference times the CPU restatement of the reference path (oracle/aaclip_oracle.py, kind "port": the
Python reference itself cannot travel to the GPU box) on the host cores, on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_IMG = 393_708_404_736          # algorithmic FLOP per image
HEAD_BYTES_IMG = 3_993_604       # fused head algorithmic bytes per image, bf16 tokens
HEAD_BYTES_IMG_F32 = 7_532_548   # fp32 tokens
METRIC = "anomaly maps/sec (336px, ViT-L-14)"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch-per-gpu", type=int, default=0, help="images per GPU and step (0: 64, or 128 at --gpus 8)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cta-group", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-drop-in", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="append the batch sweep 1..2048 (BASELINE.json configs[4])")
    ap.add_argument("--text", action="store_true", help="append the text-anchor path (BASELINE.json configs[3])")
    ap.add_argument("--surgery", action="store_true", help="append the stage-1 feature extractor of train.py (SURVEY 8(f)4)")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def kernel_source_sha() -> str:
    """sha256 over the kernel sources: ties a committed ncu capture to the code it was taken on."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "aaclip_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h")):
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def load_traffic():
    """dram bytes per launch of the dominant kernels from the ncu --set full capture committed for THIS kernel source
    (profiles/r2_ncu_traffic.json, written by tools/ncu_traffic.py); None when there is none or it is stale."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if not os.path.exists(p):
        return None, "no ncu capture committed (profiles/r2_ncu_traffic.json)"
    d = json.load(open(p))
    sha = kernel_source_sha()
    if d.get("kernel_source_sha") != sha:
        return None, (f"stale: capture taken on kernel sources {d.get('kernel_source_sha')} (commit {d.get('commit')}), "
                      f"this tree is {sha}")
    return d, f"ncu --set full, commit {d.get('commit')}, kernel sources {sha}"


class ClockSampler:
    """SM clock / throttle reasons / power every 50 ms while the timed region runs, through NVML in a thread of this
    process (a child `nvidia-smi -lms` was seen to stall single steps by ~100 ms); falls back to nvidia-smi."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.h, self.stop_flag = index, [], None, None, False
        self.t0 = self.t1 = None   # the timed region, wall clock: only samples inside it are reported
        self.sm_max, self.source = None, None

    def start(self):
        if os.environ.get("AACLIP_BENCH_NO_SAMPLER"):   # diagnostics: does the sampler perturb the timed region?
            return
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if bytes is str else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.source = "nvml"
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.h = None
        try:
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1e3
                self.rows.append((time.time(), sm, self.sm_max, pw, {n for b, n in self.REASONS.items() if bits & b}))
            except Exception:
                pass
            time.sleep(0.05)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((time.time(), float(c[0]), float(c[1]), float(c[2]),
                                  {n for n, v in zip(names, c[3:7]) if v.lower().startswith("active")}))
            except (ValueError, IndexError):
                pass

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        inside = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or r[0])]
        rows = inside if inside else self.rows[-3:]
        sm = sorted(r[1] for r in rows)
        pw = sorted(r[3] for r in rows)
        reasons = sorted(set().union(*[r[4] for r in rows])) if rows else []
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[2] for r in rows), default=None),
                "power_w": pw[len(pw) // 2] if pw else None, "reasons": reasons, "samples": len(sm),
                "source": self.source}


# ----------------------------------------------------------------------------------------------- baselines (oracle)
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import aaclip_oracle as orc
    return orc


_CPU_WEIGHTS = None


def cpu_port_rate(batch: int, n_batches: int, threads: int):
    """The oracle (CPU restatement of the reference path, test.py:80-93) on `n_batches` synthetic batches of `batch`
    images with `threads` torch threads, after one warm-up batch.  Returns (images/s, seconds)."""
    global _CPU_WEIGHTS
    import torch
    orc = _oracle()
    from aaclip_b200 import synth
    torch.set_num_threads(threads)
    cfg = synth.VIT_L_14_336
    if _CPU_WEIGHTS is None:
        _CPU_WEIGHTS = (synth.clip_state_dict(cfg, 0, text=False), synth.image_adapter_state_dict(cfg, 0))
    sd, ia = _CPU_WEIGHTS
    T = synth.anchors(cfg, 1)
    imgs = synth.images(batch * 2, cfg, seed=1)

    def one(i):
        with torch.no_grad():
            x = imgs[(i % 2) * batch:(i % 2 + 1) * batch]
            seg, det = orc.visual_forward(sd, ia, x)
            return orc.predict(seg, det, T, cfg.image_size, "Industrial")

    one(0)
    t0 = time.perf_counter()
    for i in range(n_batches):
        one(i + 1)
    dt = time.perf_counter() - t0
    return batch * n_batches / dt, dt


def cpu_baselines():
    """4 threads (the reference pins 4, test.py:28-35) and all cores, batch 1 (configs[0]) and 8."""
    cores = os.cpu_count() or 1
    legs = {}
    for name, threads, batch, n in (("all_cores_b1", cores, 1, 12), ("all_cores_b8", cores, 8, 2),
                                    ("threads4_b1", 4, 1, 4), ("threads4_b8", 4, 8, 1)):
        v, secs = cpu_port_rate(batch, n, threads)
        legs[name] = {"value": v, "unit": UNIT, "threads": threads, "batch": batch, "images": batch * n, "seconds": secs}
    head = legs["all_cores_b1"]
    return {"value": head["value"], "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{head['images']} images, batch 1 each (BASELINE.json configs[0]), {head['seconds']:.1f} s, torch fp32 "
                      "restatement of the reference path (oracle/aaclip_oracle.py) on all host cores",
            "legs": legs,
            "note": "a CPU port on host cores is context, not the comparable baseline: see torch_gpu_baseline for the "
                    "reference's own op sequence on this GPU"}


def torch_gpu_baseline(B: int, inputs, anchors, eng_maps, eng_scores):
    """the reference's op sequence (stock PyTorch / ATen / cuBLAS / cuDNN kernels, as the reference's
    model(image) + calculate_similarity_map x4 would launch them) on this B200, same weights and inputs, CUDA events.
    The oracle is the checker's restatement of that sequence; it is the baseline here, never the product."""
    import torch
    orc = _oracle()
    from aaclip_b200 import synth
    cfg = synth.VIT_L_14_336
    sd = {k: v.cuda() for k, v in synth.clip_state_dict(cfg, 0, text=False).items()}
    ia = {k: v.cuda() for k, v in synth.image_adapter_state_dict(cfg, 0).items()}
    out = {"batch": B, "what": "oracle.visual_forward + oracle.predict (model/adapter.py:67-112 + test.py:83-93) on cuda, "
                               "stock PyTorch kernels"}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)

    def run(x):
        with torch.no_grad():
            seg, det = orc.visual_forward(sd, ia, x)
            return orc.predict(seg, det, anchors, cfg.image_size, "Industrial")

    def timed(fn, reps):
        fn(inputs[0])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            r = fn(inputs[i % len(inputs)])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r

    try:
        for name, tf32, reps in (("fp32_tf32_off", False, 2), ("fp32_tf32_on", True, 3)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            ms, (m, s) = timed(run, reps)
            out[name] = {"ms_per_step": ms, "images_per_s": B / (ms / 1e3), "allow_tf32": tf32}
            if name == "fp32_tf32_off" and eng_maps is not None:   # parity of the product against it, same run
                mm = lambda x: (x - x.min()) / (x.max() - x.min())
                ref_m, ref_s = run(inputs[0])
                # ranking: pairs of images the engine orders differently from the fp32 run, and how far apart the fp32 run
                # itself puts them (random-init weights give near-tied image scores: a swap inside the score tolerance is a
                # tie, not a disagreement)
                dr = ref_s[:, None] - ref_s[None, :]
                de = eng_scores[:, None] - eng_scores[None, :]
                disc = (dr * de) < 0
                n_pairs = B * (B - 1) // 2
                n_disc = int(disc.sum().item()) // 2
                gap = float(dr.abs()[disc].max()) if n_disc else 0.0
                nm = float((mm(eng_maps) - mm(ref_m)).abs().max())
                se = float((eng_scores - ref_s).abs().max())
                out["parity_vs_engine"] = {
                    "normalised_map_max_abs": nm,
                    "raw_map_max_abs": float((eng_maps - ref_m).abs().max()),
                    "score_max_abs": se,
                    "score_spread_over_batch": float(ref_s.max() - ref_s.min()),
                    "ranking_identical": bool(torch.equal(eng_scores.argsort(), ref_s.argsort())),
                    "ranking_kendall_tau": 1.0 - 2.0 * n_disc / max(n_pairs, 1),
                    "discordant_pairs": n_disc, "pairs": n_pairs,
                    "largest_reference_gap_of_a_discordant_pair": gap,
                    "tolerance": "normalised map <= 1e-2 (north star); score <= 2e-3; ranking identical except ties: every "
                                 "discordant pair lies within the score tolerance in the fp32 run itself",
                    "pass": bool(nm <= 1e-2 and se <= 2e-3 and gap <= 2e-3)}
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True

        def run_bf16(x):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return run(x)
        ms, _ = timed(run_bf16, 5)
        out["bf16_autocast"] = {"ms_per_step": ms, "images_per_s": B / (ms / 1e3)}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    del sd, ia
    torch.cuda.empty_cache()
    return out


def run_reference(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = 2  # bounded sample of the batch-64 step
    import torch
    orc = _oracle()
    from aaclip_b200 import synth
    torch.set_num_threads(threads)
    cfg = synth.VIT_L_14_336
    sd, ia = synth.clip_state_dict(cfg, 0, text=False), synth.image_adapter_state_dict(cfg, 0)
    T = synth.anchors(cfg, 1)
    imgs = synth.images(per_step, cfg, seed=1)

    def step():
        with torch.no_grad():
            seg, det = orc.visual_forward(sd, ia, imgs)
            return orc.predict(seg, det, T, cfg.image_size, "Industrial")

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "AA-CLIP ViT-L-14-336 inference, synthetic 336x336 images, random-init weights, "
                               "visual encoder + adapters + anomaly-map head (BASELINE.json configs[1])",
                   "sample": f"{per_step} images per step of the batch-64 step", "device": "host CPU"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step}-image steps x {args.steps}, torch fp32 restatement of the reference path "
                                   "(oracle/aaclip_oracle.py) on all host cores; a CPU port, not a like-for-like baseline "
                                   "(the Python reference itself cannot travel to the GPU box)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    chatter) goes to stderr.  Returns a file object bound to the original stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


# ----------------------------------------------------------------------------------------------- extra legs
def graph_time(fn, n):
    """Device time per call of fn(i): n calls captured into ONE CUDA graph, replayed twice under CUDA events - no
    Python / launch overhead between the kernels (the entries allocate nothing and never synchronise)."""
    import torch
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(n):
                fn(i)
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        g.replay()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * n)


def head_roofline(B, cfg, anchors, peaks):
    """HBM roofline of aaclip_anomaly_head (the A7 contract: n levels of normalised tokens + det in, map + score +
    extrema out): head_stream_kernel + maps_from_dots_kernel, the kernels calculate_similarity_map /
    similarity_maps_summed run.  Inputs alternate between sets larger than the 126 MB L2."""
    import torch
    from aaclip_b200 import ops

    def rate(hb, dtype, nl):
        es = 2 if dtype == torch.bfloat16 else 4
        copies = max(2, (300 << 20) // (hb * nl * cfg.patches * cfg.embed_dim * es) + 1)
        sets = [[torch.nn.functional.normalize(torch.randn(hb, cfg.patches, cfg.embed_dim, device="cuda"), dim=-1).to(dtype)
                 for _ in range(nl)] for _ in range(copies)]
        det = torch.randn(hb, cfg.embed_dim, device="cuda")
        ms = graph_time(lambda i: ops.anomaly_head(sets[i % copies], anchors, cfg.image_size, ops.HEAD_TEST_INDUSTRIAL,
                                                   det=det, want_extrema=True), 12)
        byts = hb * (nl * cfg.patches * cfg.embed_dim * es + cfg.image_size ** 2 * 4 + cfg.embed_dim * 4 + 4 + 8)
        return {"batch": hb, "ms": ms, "achieved": byts / (ms / 1e3) / 1e9, "frac": byts / (ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                "images_per_s": hb / (ms / 1e3), "bytes_per_image": byts // hb, "input_sets": copies}

    main = rate(B, torch.bfloat16, 4)
    head = {"bound": "hbm",
            "kernel": "aaclip_anomaly_head = head_stream_kernel (cp.async.bulk-staged token stream -> one scalar per patch) + "
                      "maps_from_dots_kernel (blur, upsample, map rows, extrema), chained by programmatic dependent launch; "
                      "the kernels calculate_similarity_map / similarity_maps_summed run",
            "achieved": main["achieved"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": main["frac"], "ms": main["ms"],
            "images_per_s": main["images_per_s"], "bytes_per_image": main["bytes_per_image"],
            "how": "12 back-to-back calls captured into one CUDA graph, replayed twice under CUDA events; inputs alternate "
                   f"between {main['input_sets']} sets of 4 levels (> 126 MB L2 each pair): every call streams from HBM",
            "tokens_f32": rate(B, torch.float32, 4),
            "one_level_call_bf16": rate(B, torch.bfloat16, 1),
            "at_4x_batch": rate(4 * B, torch.bfloat16, 4)}
    return head


def drop_in_leg(B, cfg, inputs, anchors, steps):
    """The reference's own call sequence (test.py:80-93) on the drop-in classes: model(image) -> 4 levels of tokens + det,
    4 x calculate_similarity_map, cat, sum, image score - with the reference's fp32 tokens and with bf16 tokens, and the
    one-call form similarity_maps_summed.  CUDA events, device-resident inputs."""
    import torch
    from aaclip_b200 import synth
    from aaclip_b200.adapter import AdaptedCLIP
    from aaclip_b200.clip import CLIP
    from aaclip_b200.forward_utils import calculate_similarity_map, similarity_maps_summed
    clip = CLIP(cfg, text=False)
    clip.load_state_dict(synth.clip_state_dict(cfg, 0, text=False), strict=False)
    out = {"batch": B, "what": "AdaptedCLIP.forward + calculate_similarity_map x4 + cat + sum + score (test.py:80-93)"}
    for name, dt in (("tokens_f32", torch.float32), ("tokens_bf16", torch.bfloat16)):
        model = AdaptedCLIP(clip_model=clip, image_adapt_until=cfg.image_adapt_until, levels=list(cfg.levels), relu=False,
                            max_batch=B, seg_dtype=dt).to("cuda").eval()
        model.image_adapter.load_state_dict(synth.image_adapter_state_dict(cfg, 0))

        def ref_style(x):
            feats, det = model(x)
            score = ((det @ anchors)[:, 1] + 1) / 2
            maps = torch.cat([calculate_similarity_map(f, anchors, cfg.image_size, test=True, domain="Industrial")
                              for f in feats], dim=1).sum(1)
            return maps, score

        def one_call(x):
            feats, det = model(x)
            return similarity_maps_summed(feats, anchors, cfg.image_size, "Industrial", det_feature=det)

        # the two forms alternate call by call (each call bracketed by its own events): the step runs at the power cap, so
        # a form timed second would be timed on a warmer GPU
        forms = (("per_level_calls", ref_style), ("summed_call", one_call))
        for i in range(2):
            for _, fn in forms:
                fn(inputs[i % len(inputs)])
        torch.cuda.synchronize()
        evs = {lname: [] for lname, _ in forms}
        for i in range(steps):
            for lname, fn in (forms if i % 2 == 0 else forms[::-1]):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(inputs[i % len(inputs)])
                e1.record()
                evs[lname].append((e0, e1))
        torch.cuda.synchronize()
        leg = {}
        for lname, _ in forms:
            ms = sum(a.elapsed_time(b) for a, b in evs[lname]) / steps
            leg[lname] = {"ms_per_step": ms, "images_per_s": B / (ms / 1e3)}
        out[name] = leg
        model._engine.close()
        del model
        torch.cuda.empty_cache()
    return out


def sweep_leg(cfg, anchors, peaks, max_b=2048):
    """BASELINE.json configs[4]: batch 1 .. 2048 (chunks of <= 256 images), all four taps, fused head."""
    import torch
    from aaclip_b200 import synth
    from aaclip_b200.engine import Engine
    eng = Engine(cfg, device=torch.cuda.current_device(), max_batch=256, text=False)
    eng.load_state_dicts(synth.clip_state_dict(cfg, 0, text=False), synth.image_adapter_state_dict(cfg, 0), None)
    rows = []
    side = torch.cuda.Stream()
    B = 1
    while B <= max_b:
        g = torch.Generator(device="cuda").manual_seed(B)
        img = torch.randn(B, 3, cfg.image_size, cfg.image_size, device="cuda", generator=g)
        outs = (torch.empty(B, cfg.image_size, cfg.image_size, device="cuda"), torch.empty(B, device="cuda"))
        reps = max(2, min(20, 2048 // B))
        with torch.cuda.stream(side):
            for _ in range(3):   # first sight eager, second captured, third replayed (stable pointers)
                eng.forward_fused(img, anchors, out=outs)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                eng.forward_fused(img, anchors, out=outs)
            e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tf = F_IMG * B / (ms / 1e3) / 1e12
        rows.append({"batch": B, "ms": ms, "images_per_s": B / (ms / 1e3), "tflops": tf,
                     "frac_of_burst_peak": tf / peaks["bf16_tflops"], "frac_of_sustained_peak": tf / peaks["bf16_tflops_sustained"]})
        del img, outs
        B *= 2
    eng.close()
    torch.cuda.empty_cache()
    return {"workload": "BASELINE.json configs[4]: batch sweep with taps at layers 6/12/18/24, fused forward, CUDA graphs",
            "rows": rows}


def text_leg(cfg, peaks):
    """BASELINE.json configs[3]: text encoder with text adapter over the prompt templates x 15 class names (240 sentences,
    dataset/constants.py:135-147; synthetic token ids), anchors [768,2] per class, then similarity against cached patch
    features (64 images x 4 levels) for every class."""
    import torch
    from aaclip_b200 import ops, synth
    from aaclip_b200._lib import check, cur_stream, ptr
    from aaclip_b200.engine import Engine
    eng = Engine(cfg, device=torch.cuda.current_device(), max_batch=1, max_text=256, text=True)
    eng.load_state_dicts(synth.clip_state_dict(cfg, 0), synth.image_adapter_state_dict(cfg, 0), synth.text_adapter_state_dict(cfg, 0))
    n_cls, n_norm, n_abn = 15, 7, 9
    tok = synth.tokens(n_cls * (n_norm + n_abn), cfg, seed=2).cuda()
    out_a = torch.empty(n_cls, cfg.embed_dim, 2, device="cuda")

    def anchors_all():
        emb = eng.text_forward(tok)   # one batched pass over all 240 sentences
        for c in range(n_cls):
            base = c * (n_norm + n_abn)
            check(eng.lib.aaclip_text_anchor(ptr(emb[base:base + n_norm]), n_norm, cfg.t_width, ptr(out_a[c]), 0, cur_stream(emb.device)))
            check(eng.lib.aaclip_text_anchor(ptr(emb[base + n_norm:base + n_norm + n_abn]), n_abn, cfg.t_width, ptr(out_a[c]), 1,
                                             cur_stream(emb.device)))
        return out_a

    def ev(fn, reps):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms = ev(anchors_all, 5)
    A = anchors_all()
    feats = [torch.nn.functional.normalize(torch.randn(64, cfg.patches, cfg.embed_dim, device="cuda"), dim=-1).bfloat16() for _ in range(4)]
    ms2 = ev(lambda: [ops.anomaly_head(feats, A[c], cfg.image_size, ops.HEAD_TEST_INDUSTRIAL) for c in range(n_cls)], 5)
    flop = 240 * 13.6e9
    eng.close()
    torch.cuda.empty_cache()
    return {"workload": "BASELINE.json configs[3]: 240 prompt sentences x 77 tokens through the 12-layer text tower with the "
                        "text adapter -> 15 class anchors [768,2]; then 64 cached images x 4 levels against every class",
            "text_anchors_ms": ms, "sentences_per_s": 240 / (ms / 1e3), "text_tflops": flop / (ms / 1e3) / 1e12,
            "similarity_ms_15_classes": ms2, "maps_per_s": 64 * n_cls / (ms2 / 1e3),
            "similarity_hbm_frac": 64 * n_cls * HEAD_BYTES_IMG / (ms2 / 1e3) / 1e9 / peaks["hbm_gbs"]}


def surgery_leg():
    """stage 1 of train.py (lines 74-85, under no_grad) at its own defaults - 518 px, batch 2 (train.py:186,
    199), DAPM_layer 20 - and at batch 8: surgery CLIP patch features at [6,12,18,24] + the plain CLIP's class feature.
    Beside it the same op sequence in stock PyTorch on this GPU (the oracle's restatement, TF32 on and off)."""
    import torch
    orc = _oracle()
    from aaclip_b200 import synth
    from aaclip_b200.clip import CLIP
    from aaclip_b200.surgery import CLIPImageEncoder, surgery_patch_features
    cfg = synth.ModelCfg(image_size=518)
    sd = synth.clip_state_dict(cfg, 0, text=False)
    model = CLIP(cfg, text=False)
    model.load_state_dict(sd, strict=False)
    model = model.cuda()
    sd = {k: v.cuda() for k, v in sd.items()}
    levels = [6, 12, 18, 24]
    enc = CLIPImageEncoder(model, levels, surgery_until_layer=20, max_batch=8)
    plain = CLIPImageEncoder(model, [], max_batch=8)
    out = {"workload": "train.py:74-85 at 518 px (1370 tokens), DAPM_layer 20: v-v attention over the batch in the last 19 "
                       "blocks, 4 levels of projected patch features + the un-modified CLIP's class feature (2 ViT-L passes)",
           "levels": levels}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)

    def ev(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r

    try:
        for B in (2, 8):
            img = synth.images(B, cfg, seed=3).cuda()
            ms, feats = ev(lambda: surgery_patch_features(enc, plain, img), 10)
            rec = {"ms_per_batch": ms, "images_per_s": B / (ms / 1e3)}

            def ref():
                with torch.no_grad():
                    return orc.surgery_patch_features(sd, sd, img, levels=levels, surgery_until_layer=20)
            for name, tf32 in (("torch_fp32_tf32_off", False), ("torch_fp32_tf32_on", True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                rms, rf = ev(ref, 2)
                rec[name] = {"ms_per_batch": rms, "images_per_s": B / (rms / 1e3), "speedup_of_this_path": rms / ms}
                if not tf32:
                    rec["parity_vs_torch_fp32"] = {
                        "features_max_abs": max(float((a - b).abs().max()) for a, b in zip(feats, rf)),
                        "min_cosine": min(float(torch.nn.functional.cosine_similarity(a, b, dim=-1).min()) for a, b in zip(feats, rf)),
                        "tolerance": "features (two unit vectors added) <= 3e-2 max-abs, cosine >= 0.999"}
            out[f"batch_{B}"] = rec
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    del enc, plain, model, sd
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    out = _claim_stdout()
    if args.impl == "reference":
        return run_reference(args, out)

    import torch
    import torch.distributed as dist
    from aaclip_b200 import synth
    from aaclip_b200.dist import gather_scores, shard_range
    from aaclip_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = synth.VIT_L_14_336
    # BASELINE.json configs[1]: 64 images on one B200; configs[2]: 8 x B200, global batch 1024 = 128 per GPU
    B = args.batch_per_gpu if args.batch_per_gpu > 0 else (128 if world == 8 else 64)
    total = B * world
    b0, b1 = shard_range(total, rank, world)
    assert b1 - b0 == B
    peaks = load_peaks()

    # the clock sampler starts before the weights are uploaded: its NVML start-up, which can stall the driver for
    # ~100 ms, must not land in the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng = Engine(cfg, device=local_rank, max_batch=B, text=False, cta_group=args.cta_group)
    eng.load_state_dicts(synth.clip_state_dict(cfg, 0, text=False), synth.image_adapter_state_dict(cfg, 0), None)
    anchors = synth.anchors(cfg, 1).cuda()
    n_rot = 4  # rotate distinct input batches: 4 x 86.7 MB > 126 MB L2, per-step activations (>1.3 GB) exceed it anyway
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    inputs = [torch.randn(B, 3, cfg.image_size, cfg.image_size, device="cuda", generator=g) for _ in range(n_rot)]

    # caller-owned outputs, one set per rotating input: with stable pointers on a non-default stream the engine
    # replays each step as one CUDA graph launch (captured the second time a (batch, pointers) key occurs)
    outs = [(torch.empty(B, cfg.image_size, cfg.image_size, device="cuda"), torch.empty(B, device="cuda"))
            for _ in range(n_rot)]
    # N > 1: one NCCL all-gather of the B scores per step, straight into a preallocated buffer, on its own stream behind an
    # event: the collective of step i rides under the compute of step i + 1, so the ranks meet at the end of the run and
    # not at every step (a per-step rendezvous on the compute stream makes every step as slow as the slowest rank's)
    gathered = [torch.empty(total, device="cuda") for _ in range(n_rot)] if world > 1 else None
    side = torch.cuda.Stream()
    comm = torch.cuda.Stream() if world > 1 else None
    side.wait_stream(torch.cuda.current_stream())

    gather_done = [None] * n_rot   # event of the last gather that read outs[k]'s scores: the forward that rewrites them waits for it

    def step(i):
        k = i % n_rot
        if world > 1 and gather_done[k] is not None:
            torch.cuda.current_stream().wait_event(gather_done[k])   # n_rot steps old: never stalls in practice
        maps, scores = eng.forward_fused(inputs[k], anchors, "Industrial", out=outs[k])
        if world > 1:
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                scores = gather_scores(scores, total, out=gathered[k])
                gather_done[k] = torch.cuda.Event()
                gather_done[k].record()
        return maps, scores

    def sync_all():
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.cuda.stream(side):
        for i in range(2 * n_rot):     # graph priming (first sight: eager, second: capture), before the W warm-up steps
            step(i)
        for i in range(args.warmup):
            step(i)
        sync_all()
        sampler.begin()
        launches0 = eng.launch_count
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        evs[0].record()
        for i in range(args.steps):
            maps, scores = step(i)
            evs[i + 1].record()
        sync_all()
        sampler.end()
    ms = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    launches = eng.launch_count - launches0
    t = torch.tensor([ms], device="cuda")
    rank_ms = [ms / args.steps]
    if world > 1:
        # every rank's own device time, for the record: the ranks never wait for each other inside the timed region, so
        # the job's time (MAX over ranks) is the slowest GPU's, and single B200s differ by a few per cent at the power cap
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [float(x.item()) / args.steps for x in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = total * args.steps / (ms / 1e3)
    finite = bool(torch.isfinite(maps).all() and torch.isfinite(scores).all())

    # ---- per-kernel-class CUDA-event profile of the same step: events around every launch (aaclip_profile_*), run
    #      right after the timed region on the same inputs; 3 settling steps are discarded (the idle time the events
    #      insert lets the clocks rise a little above their steady state under the power cap), 3 are kept
    eng.profile(True)
    for i in range(3):
        eng.forward_fused(inputs[i % n_rot], anchors, "Industrial")
    eng.profile_read()
    prof_steps = 3
    for i in range(prof_steps):
        eng.forward_fused(inputs[i % n_rot], anchors, "Industrial")
    prof = eng.profile_read()
    eng.profile(False)
    kern = {}
    tot_ms = sum(v[0] for v in prof.values()) / prof_steps
    span_ms = eng.profile_span_ms / prof_steps   # profiled step incl. the idle time between kernels
    for name, (pms, cnt) in prof.items():
        if cnt:
            kern[name] = {"ms_per_step": pms / prof_steps, "launches_per_step": cnt // prof_steps,
                          "share": (pms / prof_steps) / tot_ms}
    gemm_ms = sum(v["ms_per_step"] for k, v in kern.items() if k.startswith("gemm_"))
    gemm_launches = sum(v["launches_per_step"] for k, v in kern.items() if k.startswith("gemm_"))
    attn_flop_img = 24 * 2 * 2 * 16 * 577 * 577 * 64
    gemm_flop_step = (F_IMG - attn_flop_img) * B
    gemm_tflops = gemm_flop_step / (gemm_ms / 1e3) / 1e12
    traffic, traffic_note = load_traffic()
    roofline = {
        "bound": "tensor", "kernel": "gemm::gemm_kernel (tcgen05, all encoder/adapter/projection GEMMs)",
        "achieved": gemm_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": gemm_tflops / peaks["bf16_tflops_sustained"], "frac_of_burst_peak": gemm_tflops / peaks["bf16_tflops"],
        "peak_source": f"{peaks['source']} cuBLAS bf16 sustained (MEASURED_PEAKS.json); burst {peaks['bf16_tflops']}",
        "flop_per_launch_avg": gemm_flop_step / max(gemm_launches, 1), "launches_per_step": gemm_launches,
        "avg_launch_ms": gemm_ms / max(gemm_launches, 1), "share_of_step": gemm_ms / tot_ms,
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the captured launches of the two largest
        # GEMM classes (c_fc and c_proj); only when the capture was taken on exactly these kernel sources
        "traffic": (None if traffic is None or B != traffic.get("batch") else traffic.get("gemm_traffic_avg_bytes")),
        "traffic_source": traffic_note,
        "traffic_detail": None if traffic is None else traffic.get("kernels"),
        "achieved_in_timed_region_estimate": gemm_tflops * (span_ms / (ms / args.steps)),
        "whole_step_tflops": F_IMG * B / (ms / args.steps / 1e3) / 1e12,
        "whole_step_frac": F_IMG * B / (ms / args.steps / 1e3) / 1e12 / peaks["bf16_tflops_sustained"],
        "whole_step_frac_of_burst_peak": F_IMG * B / (ms / args.steps / 1e3) / 1e12 / peaks["bf16_tflops"],
    }
    if "attention" in kern:
        kern["attention"]["tflops"] = attn_flop_img * B / (kern["attention"]["ms_per_step"] / 1e3) / 1e12

    # ---- e2e: same step through the host-buffer C-ABI entry (H2D of the images + D2H of maps and scores inside)
    e2e = None
    if not args.no_e2e:
        S = cfg.image_size
        h_img = [inputs[i].cpu().pin_memory() for i in range(2)]
        h_maps = [torch.empty(B, S, S).pin_memory() for _ in range(2)]
        h_scores = [torch.empty(B).pin_memory() for _ in range(2)]
        h_ext = [torch.empty(B, 2).pin_memory() for _ in range(2)]
        h_anchor = anchors.cpu()

        def e2e_loop(n):
            # the loop of test.py:get_predictions over n batches through the pipelined host-buffer entry: every
            # batch's H2D (images) and D2H (maps + scores + extrema) is inside; copies of neighbouring batches overlap compute
            prev = None
            for i in range(n):
                t = eng.submit_host(h_img[i % 2], h_anchor, h_maps[i % 2], h_scores[i % 2], extrema_out=h_ext[i % 2])
                if prev is not None:
                    eng.wait_host(prev)
                    if world > 1:
                        gather_scores(h_scores[(i - 1) % 2].cuda(non_blocking=True), total, out=gathered[(i - 1) % n_rot])
                prev = t
            eng.wait_host(prev)
            if world > 1:
                gather_scores(h_scores[(n - 1) % 2].cuda(non_blocking=True), total, out=gathered[(n - 1) % n_rot])

        e2e_loop(3)
        sync_all()
        t0 = time.perf_counter()
        e2e_loop(args.steps)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # the synchronous single-call form (H2D -> compute -> D2H serialised), for comparison
        t1 = time.perf_counter()
        for i in range(3):
            eng.forward_fused_host(h_img[i % 2], h_anchor, h_maps[0], h_scores[0])
        sync_ms = (time.perf_counter() - t1) / 3 * 1e3
        e2e = {"value": total * args.steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(h_img[0].numel() * 4 + h_anchor.numel() * 4),
               "d2h_bytes_per_step": int(h_maps[0].numel() * 4 + h_scores[0].numel() * 4 + h_ext[0].numel() * 4),
               "api": "aaclip_submit_host / aaclip_wait_host (C ABI, pinned host buffers, two batches in flight: "
                      "Engine.predict_stream); maps, scores and per-image map extrema come back",
               "synchronous_call_ms": sync_ms, "synchronous_call_images_per_s": B / (sync_ms / 1e3)}

        # the same loop fed RAW uint8 images (MVTec's 1024 x 1024 RGB): H2D of the bytes, then the loader's
        # transform_x (dataset/__init__.py:127-136: PIL bicubic resize + ToTensor + Normalize) on the device
        try:
            R = 1024
            h_raw = [torch.randint(0, 256, (B, R, R, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]

            def raw_loop(n):
                prev = None
                for i in range(n):
                    t = eng.submit_host_u8(h_raw[i % 2], h_anchor, h_maps[i % 2], h_scores[i % 2])
                    if prev is not None:
                        eng.wait_host(prev)
                    prev = t
                eng.wait_host(prev)

            raw_loop(3)
            sync_all()
            t2 = time.perf_counter()
            raw_loop(args.steps)
            torch.cuda.synchronize()
            dtr = torch.tensor([time.perf_counter() - t2], device="cuda")
            if world > 1:
                dist.all_reduce(dtr, op=dist.ReduceOp.MAX)
            e2e["from_raw_u8"] = {"value": total * args.steps / float(dtr.item()), "unit": UNIT, "raw_size": [R, R],
                                  "h2d_bytes_per_step": int(h_raw[0].numel() + h_anchor.numel() * 4),
                                  "api": "aaclip_submit_host_u8: raw RGB bytes in, transform_x on the device "
                                         "(bit-exact with PIL + torchvision), fused forward, maps + scores out"}
            del h_raw
        except Exception as e:  # never sink the headline line
            e2e["from_raw_u8"] = {"error": str(e)}

    solo = rank == 0 and world == 1
    legs = {}

    def leg(name, enabled, fn):
        if not (solo and enabled):
            return
        try:
            legs[name] = fn()
        except Exception as e:   # an extra leg must never sink the headline line
            legs[name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()

    leg("roofline_head", True, lambda: head_roofline(B, cfg, anchors, peaks))
    leg("drop_in", not args.no_drop_in, lambda: drop_in_leg(B, cfg, inputs, anchors, max(3, min(args.steps, 10))))
    leg("torch_gpu_baseline", not args.no_gpu_baseline,
        lambda: torch_gpu_baseline(B, inputs, anchors, *eng.forward_fused(inputs[0], anchors, "Industrial")))
    leg("batch_sweep", args.sweep, lambda: sweep_leg(cfg, anchors, peaks))
    leg("text_path", args.text, lambda: text_leg(cfg, peaks))
    leg("surgery_stage1", args.surgery, surgery_leg)
    if solo and "torch_gpu_baseline" in legs and "error" not in legs["torch_gpu_baseline"]:
        for k in ("fp32_tf32_off", "fp32_tf32_on", "bf16_autocast"):
            legs["torch_gpu_baseline"][k]["speedup_of_this_path"] = value / legs["torch_gpu_baseline"][k]["images_per_s"]
    leg("cpu_baseline", not args.no_cpu_baseline, cpu_baselines)

    if rank == 0:
        workload = ("single B200 bf16 batch 64: ViT-L-14-336 visual encoder + residual adapters + anomaly-map head producing "
                    "336x336 pixel maps and image scores (BASELINE.json configs[1])")
        if world > 1:
            workload = (f"{world} x B200 data-parallel, global batch {total} ({B} per GPU) synthetic images, per-GPU shards, "
                        "image-score gather over NVLink" + (" (BASELINE.json configs[2])" if total == 1024 and world == 8 else
                                                            " (BASELINE.json configs[1] per GPU)"))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload,
                       "batch_per_gpu": B, "global_batch": total, "image_size": cfg.image_size,
                       "parallelism": f"dp{world}", "weights": "random-init (synthetic, seed 0)",
                       "precision": "bf16 GEMM/attention operands, fp32 accumulation, residual stream, LayerNorm, softmax, head",
                       "l2": f"inputs rotate over {n_rot} batches ({n_rot * B * 3 * cfg.image_size ** 2 * 4 / 1e6:.0f} MB > 126 MB L2); "
                             "per-step activation traffic (>1.3 GB) exceeds L2",
                       "outputs_finite": finite},
            "roofline": roofline, "roofline_head": legs.get("roofline_head"), "kernels": kern,
            "profiled_step": {"span_ms": span_ms, "sum_kernel_ms": tot_ms, "idle_between_kernels_ms": span_ms - tot_ms,
                              "note": "3 extra steps (after 3 discarded) with CUDA events around every launch"},
            "cpu_baseline": legs.get("cpu_baseline"), "torch_gpu_baseline": legs.get("torch_gpu_baseline"),
            "drop_in": legs.get("drop_in"), "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks,
            "rank_ms_per_step": [round(x, 3) for x in rank_ms],
            "step_ms": {"min": min(step_ms), "median": sorted(step_ms)[len(step_ms) // 2], "max": max(step_ms),
                        "all": [round(x, 3) for x in step_ms]},
        }
        if "batch_sweep" in legs:
            line["batch_sweep"] = legs["batch_sweep"]
        if "text_path" in legs:
            line["text_path"] = legs["text_path"]
        if "surgery_stage1" in legs:
            line["surgery_stage1"] = legs["surgery_stage1"]
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
