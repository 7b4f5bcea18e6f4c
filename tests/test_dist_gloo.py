"""CPU, world_size 2 over gloo: the N>1 exchange steps (classifier-row all-gather, F1 count all-gather-sum,
top-k all-gather) and the class-sharded generation's equivalence to the single-process result."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from tests.helpers import O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from ovmr_b200 import dist as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    D.init_from_env(backend="gloo")
    torch.set_num_threads(2)
    cfg = O.CLIP_CONFIGS["tiny"]
    C, S = 5, 2                                       # uneven shards: 3 + 2 classes
    sd = O.init_clip_state(cfg, seed=0)
    pl = O.init_prompt_learner_state(cfg[0], n_ctx=2, seed=1)
    from ovmr_b200.clip import tokenize
    tok, vt = tokenize([f"a class {i}." for i in range(C)]), tokenize("a .")
    labels = torch.arange(C).repeat_interleave(S)
    ex = O.synth_images(C * S, 64, seed=9, structured_classes=labels)
    sh = D.class_shard(C)
    assert (sh.rank, sh.world) == (rank, world)
    with torch.no_grad():
        t_cls = O.zero_shot_classifier(sd, tok)
        # --- this rank generates only its classes (oracle arithmetic, product collectives)
        e = cfg[0]
        prompt_tokens = sd["token_embedding.weight"][tok]
        vtemp = sd["token_embedding.weight"][vt]
        my = slice(sh.lo * S, sh.hi * S)
        feats = O.l2n(O.encode_image(sd, ex[my])).reshape(sh.size, S, e)
        ex_label = torch.arange(sh.lo, sh.hi)
        mm_p, mm_l, v_p, v_l, _ = O.prompt_learner_forward(pl, prompt_tokens, vtemp, feats, ex_label,
                                                           tok[ex_label].argmax(-1))
        mm_loc, v_loc = O.get_mm_v_feats(sd, mm_p, mm_l, v_p, v_l)
        # the product's packed exchange: classifier rows + an int32 flag vector in ONE collective (uneven shards -> padded)
        flags = torch.full((sh.size,), rank + 1, dtype=torch.int32)
        mm, v, fl = D.all_gather_packed([mm_loc, v_loc, flags], C)
        assert mm.shape == (C, e) and fl.dtype == torch.int32 and fl.tolist() == [1, 1, 1, 2, 2]
        assert torch.equal(mm, D.all_gather_rows(mm_loc, C)) and torch.equal(v, D.all_gather_rows(v_loc, C))
        # --- local exemplars scored against all classifiers; count histograms summed across ranks
        scale = sd["logit_scale"].exp()
        flat = feats.reshape(sh.size * S, e)
        lab = ex_label.repeat_interleave(S)
        counts = torch.zeros(2 * C * 3 + C, dtype=torch.int32)
        for s_i, w in enumerate((mm, v, t_cls)):
            pred = (scale * flat @ w.t()).argmax(1)
            for p_, y_ in zip(pred.tolist(), lab.tolist()):
                counts[C * 3 + p_ * 3 + s_i] += 1
                if p_ == y_:
                    counts[y_ * 3 + s_i] += 1
        for y_ in lab.tolist():
            counts[2 * C * 3 + y_] += 1
        counts = D.all_gather_sum(counts)
        tp = counts[:C * 3].view(C, 3).float()
        npred = counts[C * 3:2 * C * 3].view(C, 3).float()
        nlab = counts[2 * C * 3:].float().unsqueeze(1)
        p, r = tp / npred, tp / nlab
        fw = (10.0 * torch.nan_to_num(2 * p * r / (p + r))).softmax(-1)
        # --- query shards + top-k all-gather
        Q = 7
        qs = O.synth_images(Q, 64, seed=10)
        lo, hi = D.shard_range(Q, rank, world)
        probs = O.classify(scale, O.l2n(O.encode_image(sd, qs[lo:hi])),
                           {"mm_classifier": mm, "vision_classifier": v, "text_classifier": t_cls, "fusion_weight": fw})
        idx, val = O.topk(probs, 1)
        idx_all, val_all = D.all_gather_packed([idx.int(), val], Q)       # one collective for (index, probability)
        # equal shards take the zero-copy path
        eq = D.all_gather_rows(torch.full((3, 2), float(rank)), 6)
        assert eq.shape == (6, 2) and eq[:3].eq(0).all() and eq[3:].eq(1).all()
        elapsed = D.max_over_ranks(float(rank + 1), torch.device("cpu"))
        assert elapsed == float(world)
    D.barrier()
    if rank == 0:
        torch.save({"mm": mm, "v": v, "fw": fw, "idx": idx_all, "val": val_all}, os.path.join(out_dir, "sharded.pt"))
    dist.destroy_process_group()


def test_sharded_generation_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = torch.load(tmp_path / "sharded.pt")
    cfg = O.CLIP_CONFIGS["tiny"]
    C, S, Q = 5, 2, 7
    sd = O.init_clip_state(cfg, seed=0)
    pl = O.init_prompt_learner_state(cfg[0], n_ctx=2, seed=1)
    from ovmr_b200.clip import tokenize
    tok, vt = tokenize([f"a class {i}." for i in range(C)]), tokenize("a .")
    labels = torch.arange(C).repeat_interleave(S)
    ex = O.synth_images(C * S, 64, seed=9, structured_classes=labels)
    with torch.no_grad():
        t_cls = O.zero_shot_classifier(sd, tok)
        gen = O.forward_prompt(sd, pl, tok, vt, t_cls, [(ex, labels)], S, tau=10.0)
        probs = O.classify(sd["logit_scale"].exp(), O.l2n(O.encode_image(sd, O.synth_images(Q, 64, seed=10))), gen)
        idx, val = O.topk(probs, 1)
    assert (got["mm"] - gen["mm_classifier"]).abs().max() < 2e-6
    assert (got["v"] - gen["vision_classifier"]).abs().max() < 2e-6
    assert (got["fw"] - gen["fusion_weight"]).abs().max() < 1e-6
    assert torch.equal(got["idx"].long(), idx) and (got["val"] - val).abs().max() < 2e-6
