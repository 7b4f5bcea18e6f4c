"""CPU: the C-ABI library loads and exports every symbol include/ovmr_b200.h declares; host-side logic
(config, sharding, loaders, error behaviour) — no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from tests.helpers import O, ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ovmr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ovmr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ovmr_b200 import _lib as L
    lib = L.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ovmr_b200.h but not exported"
    assert sorted(L.SIGNATURES) == declared, "ctypes SIGNATURES out of sync with the header"
    assert lib.ovmr_abi_version() == 2
    assert isinstance(lib.ovmr_last_error(), bytes)


def test_abi_struct_layout_matches_header():
    from ovmr_b200 import _lib as L
    p = ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(L.BlockWeights) == 18 * p   # 12 operands + 6 LayerNorm-folded operands
    assert ctypes.sizeof(L.Transformer) == 4 * 4 + p
    assert L.Vit.transformer.offset == 5 * 4 + 4 + 8 * p  # 5 ints, padding, 8 pointers
    assert L.Text.transformer.offset == 3 * 4 + 4 + 4 * p


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly (never fall back to torch / the oracle)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ovmr_b200 import _lib as L
    from ovmr_b200.clip.model import CLIP
    with pytest.raises(L.OvmrNativeError):
        L.lib()
    m = CLIP(*O.CLIP_CONFIGS["tiny"]).eval()
    with pytest.raises(L.OvmrNativeError):
        m.encode_image(torch.zeros(1, 3, 64, 64))
    with pytest.raises(L.OvmrNativeError):
        m.encode_text(torch.zeros(1, 77, dtype=torch.long))
    with pytest.raises(L.OvmrNativeError):
        m.ln_final(torch.zeros(2, 128))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ovmr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("Oracle", ""), f"{f} references the oracle"


def test_state_dict_contract_and_arch_inference():
    from ovmr_b200.clip.model import CLIP, _arch_from_state_dict, build_model, build_model_fp32
    cfg = O.CLIP_CONFIGS["tiny"]
    sd = O.init_clip_state(cfg, seed=0)
    m = CLIP(*cfg)
    assert set(m.state_dict()) == set(sd)
    for k, v in m.state_dict().items():
        assert v.shape == sd[k].shape, k
    assert _arch_from_state_dict(sd) == cfg
    m16 = build_model(sd)
    assert m16.dtype == torch.float16 and not m16.training        # clip/model.py:899-936 converts to fp16
    assert build_model_fp32(sd).dtype == torch.float32
    assert _arch_from_state_dict(O.init_clip_state((64 * 2, 32, 1, 128, 16, 77, 49408, 128, 2, 1))) == \
        (128, 32, 1, 128, 16, 77, 49408, 128, 2, 1)
    with pytest.raises(NotImplementedError):
        CLIP(1024, 224, (3, 4, 6, 3), 64, None, 77, 49408, 512, 8, 12)   # ModifiedResNet: out of scope


def test_clip_load_and_tokenize_errors(tmp_path):
    from ovmr_b200 import clip
    assert "ViT-B/16" in clip.available_models() and "ViT-L/14@336px" in clip.available_models()
    assert len(clip.available_models()) == 8
    with pytest.raises(RuntimeError):
        clip.load("no-such-model")
    with pytest.raises(RuntimeError):
        clip.tokenize("word " * 100)
    t = clip.tokenize("word " * 100, truncate=True)
    assert t.shape == (1, 77) and int(t[0, -1]) == 49407
    assert clip.tokenize(["a", "b c"]).dtype == torch.long
    # offline entry: a state_dict file
    path = tmp_path / "tiny.pt"
    torch.save(O.init_clip_state(O.CLIP_CONFIGS["tiny"]), path)
    with pytest.warns(UserWarning):
        model, preprocess = clip.load(str(path), device="cpu")
    assert model.visual.input_resolution == 64 and model.dtype == torch.float32 and callable(preprocess)


def test_precision_policy(monkeypatch):
    from ovmr_b200 import config
    p = config.Precision("mixed")
    assert (p.vision_fp16, p.text_fp16) == (False, True)
    assert (config.Precision("bf16").vision_fp16, config.Precision("bf16").text_fp16) == (False, False)
    assert (config.Precision("fp16").vision_fp16, config.Precision("fp16").text_fp16) == (True, True)
    with pytest.raises(ValueError):
        config.Precision("int8")
    cfg = config.make_cfg(shots=4)
    assert cfg.DATASET.NUM_SHOTS == 4 and cfg.TRAINER.COCOOP.N_CTX == 2 and cfg.EVAL_MODE == "fusion"


def test_shard_ranges_partition_exactly():
    from ovmr_b200.dist import shard_range
    for n in (0, 1, 7, 1000, 21841, 50000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_plan_batches_partitions_exactly_and_respects_cap():
    from ovmr_b200.data import plan_batches
    for n, cap, unit in ((1000, 32, 16), (50000, 512, 1), (125, 32, 16), (6250, 512, 1), (1, 512, 1), (513, 512, 1),
                         (21841, 128, 4), (0, 512, 1)):
        plan = plan_batches(n, cap, unit)
        assert sum(z for _, z in plan) == n
        assert all(0 < z <= cap for _, z in plan)
        off = 0
        for o, z in plan:
            assert o == off
            off += z


def test_plan_batches_wave_model_choices():
    """The planner compares 'full batches + ragged tail' with 'even split' under a wave model of the persistent GEMMs: it must
    never pick the costlier candidate, must keep a 512-image ViT-B/16 batch whole (394 row blocks = 47.9 / 63.9 / 17.9 waves:
    already within 0.5 % of whole waves), and must behave for the ViT-L geometry (577 tokens, width 1024: residual GEMMs on the
    CTA pairs, not on row-block clusters)."""
    from ovmr_b200.data import plan_batches

    def waves(z, tokens, width):
        mp = -(-z * tokens // 256)
        nt = width // 256
        c = -(-mp * 3 * nt // 74) + -(-mp * 4 * nt // 74)
        c += (-(-mp // 22) if width <= 768 else -(-mp * nt // 74)) * 5
        return c

    for n, cap, tokens, width in ((8250, 512, 197, 768), (2000, 512, 197, 768), (6250, 256, 577, 1024), (16000, 512, 197, 768),
                                  (87364, 512, 197, 768), (777, 256, 577, 1024)):
        plan = plan_batches(n, cap, tokens_per_image=tokens, width=width)
        sizes = [z for _, z in plan]
        k = -(-n // cap)
        even = [n // k + (1 if i < n % k else 0) for i in range(k)]
        ragged = [cap] * (n // cap) + ([n % cap] if n % cap else [])
        cost = lambda zs: sum(waves(z, tokens, width) for z in zs)
        assert sizes in (even, ragged)
        assert cost(sizes) == min(cost(even), cost(ragged))
    assert [z for _, z in plan_batches(1024, 512)] == [512, 512]


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's own CPU implementation on the host cores — its unmodified code when
    /root/reference or oracle/_ref is present, else the oracle port): one JSON line with the contract keys, labelled
    with the shape that was actually run."""
    import json
    import subprocess
    import sys
    from oracle import ref_loader as R
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample-images", "4", "--shots", "2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "img/s" and line["value"] > 0 and line["higher_is_better"] is True
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["cpu_baseline"]["kind"] == ("reference" if R.reference_available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    # --shots 2 is not a BASELINE shape: metric and workload strings must say so instead of claiming config 2
    assert "x 2 shot" in line["metric"] and line["config"]["workload"].startswith("CUSTOM shape")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0



def test_u8_normalization_exactness_check_matches_exact_rational_arithmetic():
    """The host check that licenses the division-free ToTensor + Normalize in the uint8 kernels (q = a r; q += fma(-b, q, a) r):
    its verdict for CLIP's mean / std (clip/clip.py:79) and for a spread of other constants must equal an independent
    evaluation of the same formula in exact rational arithmetic with correct rounding to fp32 (no GPU involved)."""
    import ctypes as C
    from fractions import Fraction as F

    import numpy as np
    from ovmr_b200 import _lib as L
    lib = L.load()

    def rn(x):      # round a Fraction to the nearest float32 (ties to even via numpy on the two neighbours)
        y = np.float32(float(x))
        cands = [y, np.nextafter(y, np.float32(np.inf)), np.nextafter(y, np.float32(-np.inf))]
        return min(cands, key=lambda c: (abs(F(float(c)) - x), int(np.float32(c).view(np.uint32)) & 1))

    def fma(a, b, c):
        return rn(F(float(a)) * F(float(b)) + F(float(c)))

    def exact(mean_std):
        f32 = np.float32
        for c in range(3):
            m, sd = f32(mean_std[c]), f32(mean_std[3 + c])
            rsd, r255 = f32(1.0) / sd, f32(1.0) / f32(255.0)
            for v in range(256):
                f = f32(v)
                ref = ((f / f32(255.0)) - m) / sd
                t = f * r255
                t = fma(fma(f32(-255.0), t, f), r255, t)
                u = t - m
                y = u * rsd
                y = fma(fma(-sd, y, u), rsd, y)
                if y != ref:
                    return 0
        return 1

    clip_ms = [0.48145466, 0.4578275, 0.40821073, 0.26862954, 0.26130258, 0.27577711]
    assert lib.ovmr_u8_normalization_is_exact((C.c_float * 6)(*clip_ms)) == 1 == exact(clip_ms)
    rng = np.random.default_rng(0)
    for _ in range(6):
        ms = list(rng.uniform(0.2, 0.6, 3)) + list(rng.uniform(0.05, 0.6, 3))
        assert lib.ovmr_u8_normalization_is_exact((C.c_float * 6)(*ms)) == exact(ms), ms
    assert lib.ovmr_u8_normalization_is_exact((C.c_float * 6)(0.5, 0.5, 0.5, 0.0, 0.2, 0.2)) == 0     # std must be positive


def test_head_routing_by_row_count(monkeypatch):
    """ovmr_b200.engine picks the one-kernel head from FUSED_HEAD_MIN_ROWS feature rows on, the explicit head below;
    OVMR_FUSED_HEAD forces one form."""
    from ovmr_b200 import engine as Eng
    monkeypatch.delenv("OVMR_FUSED_HEAD", raising=False)
    assert not Eng.fused_head_enabled(512) and not Eng.fused_head_enabled(Eng.FUSED_HEAD_MIN_ROWS - 1)
    assert Eng.fused_head_enabled(Eng.FUSED_HEAD_MIN_ROWS) and Eng.fused_head_enabled(87364)
    monkeypatch.setenv("OVMR_FUSED_HEAD", "1")
    assert Eng.fused_head_enabled(1)
    monkeypatch.setenv("OVMR_FUSED_HEAD", "0")
    assert not Eng.fused_head_enabled(10 ** 6)
