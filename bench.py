#!/usr/bin/env python
"""bench.py — OVMR hot path on B200: exemplar img/s + query img/s (ViT-B/16 fusion).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU fp32 port of the reference

Workload (BASELINE.json configs[1]): ViT-B/16, fusion eval, 1000 classes x 16-shot synthetic exemplars +
50,000 synthetic 224^2 queries, random-init weights (bf16-representable), EVAL_TAU 10, n_ctx 2.
One STEP = one full pass of the hot path over that workload: classifier generation for all classes
(exemplar encoding -> visual tokens -> mm / v / t classifiers -> F1 fusion weights) followed by fused
classification + top-1 of all queries.  With N > 1 the same total work is sharded (classes for generation,
queries for classification; "strong" scaling) and the only collectives are the all-gathers of the classifier
rows / F1 counts and of the top-k results.

Printed JSON (one line, rank 0): see the contract in the task statement; `value` is timed with inputs
resident in HBM, `e2e` through the public API (CustomCLIP.forward_prompt / predict_topk) with pinned-host
inputs copied every step and the results read back; `roofline` is the GEMM kernel class measured live with
CUDA events inside the timed region; `cpu_baseline` is the oracle port timed on the host cores.
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

VITB16 = (512, 224, 12, 768, 16, 77, 49408, 512, 8, 12)
GFLOP_PER_IMAGE = 35.127           # SURVEY.md §8d (ViT-B/16: GEMM 33.70 + attention 1.43)
METRIC = "exemplar+query img/s (ViT-B/16 fusion, 1000 cls x 16 shot + 50k queries)"
UNIT = "img/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--shots", type=int, default=16)
    ap.add_argument("--queries", type=int, default=50000)
    ap.add_argument("--batch", type=int, default=512, help="images per encoder call (TEST.BATCH_SIZE of the reference)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-classes", type=int, default=12,
                    help="classes in the bounded CPU sample (x shots exemplars + as many queries): ~10-30 s of CPU work")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 9:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if r[3].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s, p in zip(sm, pw)] if not pw else [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": sorted(reasons)}


def measured_traffic():
    """DRAM bytes per GEMM launch (mean over the GEMM class) from the committed ncu launch list of this command."""
    p = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("dram_bytes_per_launch"), f"ncu dram__bytes_read+write per launch, mean over {d.get('launches')} GEMM launches ({d.get('source')})"
    return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "hbm_gbs": d.get("hbm_gbs"),
                "source": "measured (MEASURED_PEAKS.json, bf16_tflops_sustained)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (torch CPU fp32) on a bounded sample
# ----------------------------------------------------------------------------------------------
_CPU_CTX = {}


def cpu_port_step(n_classes: int, shots: int, n_queries: int):
    """One bounded sample of the workload through the CPU port of the reference: generation for n_classes
    (shots exemplars each) + fused classification of n_queries.  Returns seconds."""
    from oracle import ovmr_oracle as O
    from ovmr_b200.clip import tokenize
    if not _CPU_CTX:
        torch.set_num_threads(os.cpu_count() or 1)
        _CPU_CTX["sd"] = O.init_clip_state(VITB16, seed=0)
        _CPU_CTX["pl"] = O.init_prompt_learner_state(512, n_ctx=2, seed=1)
    sd, pl = _CPU_CTX["sd"], _CPU_CTX["pl"]
    labels = torch.arange(n_classes).repeat_interleave(shots)
    ex = O.synth_images(n_classes * shots, 224, seed=1)
    qs = O.synth_images(n_queries, 224, seed=1001)
    tok = tokenize([f"a class {i}." for i in range(n_classes)])
    vt = tokenize("a .")
    t0 = time.perf_counter()
    with torch.no_grad():
        t_cls = O.zero_shot_classifier(sd, tok)
        gen = O.forward_prompt(sd, pl, tok, vt, t_cls, [(ex, labels)], shots, tau=10.0)
        qf = O.l2n(O.encode_image(sd, qs))
        probs = O.classify(sd["logit_scale"].exp(), qf, gen, "fusion")
        O.topk(probs, 1)
    return time.perf_counter() - t0


def run_reference(args, rank):
    if rank != 0:
        return
    nc, s = args.cpu_sample_classes, args.shots
    nq = nc * s
    n_img = nc * s + nq
    for _ in range(max(1, min(args.warmup, 1))):   # 1 warm-up is enough on CPU (no autotuning, no lazy init)
        cpu_port_step(nc, s, nq)
    times = [cpu_port_step(nc, s, nq) for _ in range(max(1, args.steps))]
    t = sum(times) / len(times)
    v = n_img / t
    sample = f"{nc} classes x {s} shots generation + {nq} queries (fusion, top-1) per step, ViT-B/16 fp32"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: ViT-B/16 fusion, 1000 cls x 16 shot + 50k queries",
                       "note": "CPU port of the reference (oracle/ovmr_oracle.py, pinned to the reference's outputs); "
                               "bounded sample per step, throughput extrapolates linearly (encoder-bound)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------
def build_model(args, device):
    from ovmr_b200.clip.model import CLIP
    from ovmr_b200.config import make_cfg
    from ovmr_b200.trainers.mm_classifier_one_prompt import CustomCLIP
    torch.manual_seed(0)
    clip_model = CLIP(*VITB16).eval()
    with torch.no_grad():
        for p in clip_model.parameters():
            p.copy_(p.bfloat16().float())
    clip_model = clip_model.to(device)
    cfg = make_cfg(n_ctx=2, shots=args.shots, image_size=224, eval_mode="fusion", eval_tau=10, output_dir=None)
    torch.manual_seed(1)
    model = CustomCLIP(cfg, [f"class_{i}" for i in range(args.classes)], clip_model).eval()
    return model


def device_images(n, device, seed):
    """fp32 N(0,1) images generated on the device in chunks (resident-in-HBM inputs for `value`)."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty(n, 3, 224, 224, dtype=torch.float32, device=device)
    for i in range(0, n, 1024):
        out[i:i + 1024].normal_(generator=g)
    return out


def main():
    args = parse_args()
    from ovmr_b200 import dist as D
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, rank)
        return
    rank, local_rank, world = D.init_from_env()
    assert torch.cuda.is_available(), "bench.py (impl ours) needs a GPU; there is no CPU fallback"
    device = torch.device("cuda", local_rank)
    from ovmr_b200 import _lib as L
    from ovmr_b200.data import DevicePrefetcher, plan_batches
    from ovmr_b200.config import precision

    C, S, Q, B = args.classes, args.shots, args.queries, args.batch
    cls_per_batch = max(1, B // S)
    model = build_model(args, device)
    model.image_encoder.engine(device).max_batch = max(B, 1)   # images per tower call = the bench batch
    shard = D.class_shard(C, rank, world)
    q_lo, q_hi = D.shard_range(Q, rank, world)
    n_ex_local, n_q_local = shard.size * S, q_hi - q_lo

    # ---- inputs resident in HBM
    ex_dev = device_images(n_ex_local, device, seed=1 + rank)
    q_dev = device_images(n_q_local, device, seed=1001 + rank)
    ex_labels = torch.arange(shard.lo, shard.hi, device=device).repeat_interleave(S)

    ex_plan = [(o * S, z * S) for o, z in plan_batches(shard.size, cls_per_batch, unit=S)]   # whole classes per batch
    q_plan = plan_batches(n_q_local, B)

    def exemplar_batches(images, labels):
        return [{"img": images[o:o + z], "label": labels[o:o + z]} for o, z in ex_plan]

    def query_batches(images):
        return [images[o:o + z] for o, z in q_plan]

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def one_step(ex_loader, q_iter, phase_events=None):
        model.mm_classifier = None
        if phase_events is not None:
            phase_events[0].record()
        model.forward_prompt(ex_loader, shard=shard)
        if phase_events is not None:
            phase_events[1].record()
        idxs, vals = [], []
        for qb in q_iter:
            img = qb["img"] if isinstance(qb, dict) else qb
            i, v = model.predict_topk(img, k=1)
            idxs.append(i)
            vals.append(v)
        idx = torch.cat(idxs) if idxs else torch.empty(0, 1, dtype=torch.int32, device=device)
        val = torch.cat(vals) if vals else torch.empty(0, 1, device=device)
        idx_all = D.all_gather_rows(idx, Q)
        val_all = D.all_gather_rows(val, Q)
        if phase_events is not None:
            phase_events[2].record()
        return idx_all, val_all

    dev_ex_loader = exemplar_batches(ex_dev, ex_labels)
    dev_q_batches = query_batches(q_dev)

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            one_step(dev_ex_loader, dev_q_batches)
        torch.cuda.synchronize()

        # ---- timed region 1: inputs resident in HBM (value).  No per-launch events here: the step is timed as the
        #      user would run it (event records between launches also defeat programmatic dependent launch).
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        D.barrier()
        torch.cuda.synchronize()
        launches0 = L.launch_count()
        e0, e1 = ev(), ev()
        phases = [[ev(), ev(), ev()] for _ in range(args.steps)]
        e0.record()
        for k in range(args.steps):
            one_step(dev_ex_loader, dev_q_batches, phases[k])
        e1.record()
        torch.cuda.synchronize()
        D.barrier()
        ms_total = e0.elapsed_time(e1)
        launches = L.launch_count() - launches0
        ms_step = D.max_over_ranks(ms_total / args.steps, device)
        gen_ms = D.max_over_ranks(sum(p[0].elapsed_time(p[1]) for p in phases) / args.steps, device)
        cls_ms = D.max_over_ranks(sum(p[1].elapsed_time(p[2]) for p in phases) / args.steps, device)

        # ---- timed region 1b: the same K steps again with every launch bracketed by CUDA events on its stream
        #      (ovmr_profile_*): per-kernel-class device time and algorithmic work for the roofline leg.
        D.barrier()
        torch.cuda.synchronize()
        L.profile_enable(True)
        p0, p1 = ev(), ev()
        p0.record()
        for k in range(args.steps):
            one_step(dev_ex_loader, dev_q_batches)
        p1.record()
        torch.cuda.synchronize()
        prof_ms_step = p0.elapsed_time(p1) / args.steps
        prof = L.profile_summary()
        L.profile_enable(False)
        clocks = sampler.stop() if rank == 0 else None
        D.barrier()

        # ---- timed region 2: end to end through the public API, pinned host inputs, results read back
        e2e = None
        if not args.no_e2e:
            pool_n = 4
            g = torch.Generator().manual_seed(7 + rank)

            def run_e2e(pool, what):
                def host_ex_loader():
                    for bi, (o, z) in enumerate(ex_plan):
                        yield {"img": pool[bi % pool_n][:z], "label": ex_labels[o:o + z]}

                def host_q_loader():
                    for bi, (o, z) in enumerate(q_plan):
                        yield {"img": pool[bi % pool_n][:z]}

                pf_e = DevicePrefetcher((), device)    # staging rings are allocated once and reused every step
                pf_q = DevicePrefetcher((), device)

                def e2e_step():
                    pf_e.batches, pf_q.batches = host_ex_loader(), host_q_loader()
                    b0 = pf_e.h2d_bytes + pf_q.h2d_bytes
                    idx_all, val_all = one_step(pf_e, pf_q)
                    res = (idx_all.cpu(), val_all.cpu(), model.fusion_weight.cpu())   # device -> host read of the results
                    return pf_e.h2d_bytes + pf_q.h2d_bytes - b0, sum(t.numel() * t.element_size() for t in res)

                e2e_step()
                torch.cuda.synchronize()
                D.barrier()
                a0, a1 = ev(), ev()
                a0.record()
                for _ in range(args.steps):
                    h2d, d2h = e2e_step()
                a1.record()
                torch.cuda.synchronize()
                D.barrier()
                ms = D.max_over_ranks(a0.elapsed_time(a1) / args.steps, device)
                return {"value": (C * S + Q) / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": ms, "input": what}

            # (a) uint8 pixels (what a decoder / crop produces): ToTensor + Normalize run fused on the GPU
            pool_u8 = [torch.randint(0, 256, (B, 3, 224, 224), generator=g, dtype=torch.uint8).pin_memory()
                       for _ in range(pool_n)]
            e2e = run_e2e(pool_u8, "uint8 NCHW pixels, ToTensor+Normalize fused into the patch load")
            del pool_u8
            # (b) fp32 tensors as the reference's CPU transform hands them over (4x the H2D bytes)
            pool_f32 = [torch.randn(B, 3, 224, 224, generator=g).pin_memory() for _ in range(pool_n)]
            e2e_f32 = run_e2e(pool_f32, "fp32 NCHW tensors (already normalised on the host, as the reference's DataLoader)")
            del pool_f32
            e2e["api"] = ("CustomCLIP.forward_prompt(loader) + CustomCLIP.predict_topk(images); pinned host batches "
                          "staged by ovmr_b200.data.DevicePrefetcher (per-rank bytes)")
            e2e["fp32_input"] = {k: e2e_f32[k] for k in ("value", "h2d_bytes_per_step", "ms_per_step", "input")}

    if rank != 0:
        return
    peaks = measured_peaks()
    traffic, traffic_src = measured_traffic()
    gemm = prof["gemm"]
    achieved = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
    kernel_ms = {k: round(v["ms"] / args.steps, 3) for k, v in prof.items()}
    line = {
        "metric": METRIC, "value": (C * S + Q) / (ms_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16" if not precision().vision_fp16 else "fp16", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: ViT-B/16 fusion eval, {C} classes x {S}-shot exemplars + {Q} queries "
                               f"224^2, random-init bf16-representable weights, n_ctx 2, EVAL_TAU 10",
                   "batch": B, "parallelism": f"dp{world}: classes sharded for generation, queries for classification",
                   "precision": f"{precision().mode}: image encoder {'fp16' if precision().vision_fp16 else 'bf16'} "
                                f"operands, text/aggregator towers {'fp16' if precision().text_fp16 else 'bf16'}, "
                                f"fp32 accumulate/residual/statistics",
                   "l2_policy": f"inputs larger than L2: {(n_ex_local + n_q_local) * 602112 / 1e9:.1f} GB of images per "
                                f"rank per step, ~{B * 3.6e-3:.1f} GB of activations per batch (L2 = 126 MB)"},
        "exemplar_img_s": C * S / (gen_ms / 1e3), "query_img_s": Q / (cls_ms / 1e3),
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "gemm_tn_kernel (tcgen05/TMEM GEMM: QKV, out-proj, MLP, patch-embed, projections, logits)",
                     "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops"] if peaks["tflops"] else None, "traffic": traffic,
                     "traffic_source": traffic_src,
                     "peak_source": peaks["source"], "launches_per_step": gemm["launches"] // max(1, args.steps),
                     "kernel_ms_per_step": kernel_ms, "ms_per_step_with_events": prof_ms_step,
                     "end_to_end_tensor_frac": ((C * S + Q) / world * GFLOP_PER_IMAGE / 1e3) / (ms_step / 1e3) / peaks["tflops"]},
    }
    # the other kernel classes against their own rooflines (same event-timed pass): HBM-bound classes in GB/s of
    # algorithmic bytes against the measured copy bandwidth, attention in TFLOP/s of algorithmic FLOPs
    def _cls(name, bound, unit_scale, peak):
        c = prof[name]
        ach = c["work"] / (c["ms"] * 1e-3) / unit_scale if c["ms"] > 0 else 0.0
        return {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                "frac": ach / peak if peak else None, "launches_per_step": c["launches"] // max(1, args.steps)}
    line["roofline"]["other_classes"] = [
        _cls("layernorm", "hbm", 1e9, peaks["hbm_gbs"]),
        _cls("patchify", "hbm", 1e9, peaks["hbm_gbs"]),
        _cls("head", "hbm", 1e9, peaks["hbm_gbs"]),
        _cls("attention", "tensor", 1e12, peaks["tflops"]),
    ]
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        nc, s = args.cpu_sample_classes, S
        cpu_port_step(1, s, s)  # warm-up (weight init, thread pool)
        t = cpu_port_step(nc, s, nc * s)
        line["cpu_baseline"] = {"value": (2 * nc * s) / t, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{nc} classes x {s} shots generation + {nc * s} queries (fusion, top-1), "
                                          f"ViT-B/16 fp32 oracle port, {t:.1f} s"}
    print(json.dumps(line))


if __name__ == "__main__":
    try:
        main()
    finally:
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized():
            _dist.destroy_process_group()
