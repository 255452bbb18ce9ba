"""CPU: the dependency levelling the wavefront sampler relies on (ps_lmconv_levels_host, csrc/lmconv_tc.cu) and the
weight schedule the host packs for it (pixelsynth_b200/lmconv.py)."""
import os
import sys

import numpy as np
import torch

from util import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _cases():
    import make_lmconv_golden as mk
    import pixelsynth_b200.lmconv as lm

    bgs = mk.background_cases()
    half = torch.zeros(1, 256, 256, dtype=torch.bool)
    half[:, :, 128:] = True                                   # BASELINE configs[2]: right half of the code grid masked
    bgs = torch.cat([bgs, half])
    _, order, words, smask = lm.glue_host(bgs)
    return lm, order, words, smask.reshape(len(order), -1)


def _decode(rows):
    r = rows.view(np.uint32).reshape(-1, 4)
    return r[:, 0] >> 10, r[:, 0] & 1023, r[:, 1] & 0xffff, r[:, 1] >> 16, r[:, 2] & 0xffff, r[:, 2] >> 16, r[:, 3]


def test_levels_are_wavefronts_of_independent_cells():
    lm, order, words, smask = _cases()
    rows, offs, first_b = lm.LmconvB200.levels_host(order, words, smask, 0)
    b, cell, w0, w1, w2, flags, uidx = _decode(rows)
    B = len(order)
    assert 0 < first_b < len(offs) - 1
    level = np.full((B, 1024), -1)
    for l in range(len(offs) - 1):
        level[b[offs[l]:offs[l + 1]], cell[offs[l]:offs[l + 1]]] = l
    rank = np.empty((B, 1024), int)
    for i in range(B):
        rank[i, order[i]] = np.arange(1024)
    for i in range(B):
        last = max(rank[i, c] for c in range(1024) if smask[i, c])
        for c in range(1024):
            # exactly the cells up to the last sampled one are scheduled
            assert (level[i, c] >= 0) == (rank[i, c] <= last)
            if level[i, c] < 0:
                continue
            deps = []
            for m, dil in ((0, 1), (1, 1), (2, 2)):
                for t in range(9):
                    if t != 4 and (words[i, m, c] >> t) & 1:
                        deps.append(c + ((t // 3 - 1) * 32 + (t % 3 - 1)) * dil)
            # every masked-in neighbour sits on a strictly lower level
            assert all(0 <= level[i, d] < level[i, c] for d in deps)
            # phase B = sampled cells and whatever has one among its ancestors; the levels before first_b are the
            # known prefix.  Within its phase a cell sits directly above the highest of its same-phase neighbours.
            in_b = bool(smask[i, c]) or any(level[i, d] >= first_b for d in deps)
            assert (level[i, c] >= first_b) == in_b
            if in_b:
                below = [level[i, d] for d in deps if level[i, d] >= first_b]
                assert level[i, c] == (max(below) + 1 if below else first_b)
            else:
                assert level[i, c] == (max(level[i, d] for d in deps) + 1 if deps else 0)
    # rows carry the masks, the sampled flag and the index of their uniform number (rank among the sampled cells)
    assert np.array_equal(w0, words[b, 0, cell]) and np.array_equal(w1, words[b, 1, cell]) and np.array_equal(w2, words[b, 2, cell])
    assert np.array_equal((flags & 1).astype(bool), smask[b, cell]) and np.all(flags & 4)
    for i in range(B):
        seq = [c for c in order[i] if smask[i, c]]
        got = {int(c): int(u) for bb, c, u, f in zip(b, cell, uidx, flags) if bb == i and f & 1}
        assert got == {int(c): k for k, c in enumerate(seq)}
    # the half-plane case of BASELINE configs[2]: 512 sampled cells in < 100 wavefronts instead of 512 serial steps
    n_half = level[B - 1].max() + 1 - first_b
    assert smask[B - 1].sum() == 512 and n_half < 100


def test_logits_mode_levels_every_cell():
    lm, order, words, smask = _cases()
    rows, offs, first_b = lm.LmconvB200.levels_host(order[:2], words[:2], None, 1)
    b, cell, w0, w1, w2, flags, uidx = _decode(rows)
    assert first_b == len(offs) - 1          # teacher forcing: every level is prefix
    assert len(rows) == 2048 and np.all(flags & 2) and not np.any(flags & 1)
    assert sorted(zip(b.tolist(), cell.tolist())) == [(i, c) for i in range(2) for c in range(1024)]


def test_images_with_nothing_to_sample_produce_no_rows():
    lm, order, words, smask = _cases()
    rows, offs, first_b = lm.LmconvB200.levels_host(order[:2], words[:2], np.zeros((2, 1024), np.uint8), 0)
    assert len(rows) == 0 and len(offs) == 1 and first_b == 0


def test_weight_schedule_tiles_reproduce_the_layers():
    """Un-swizzling the packed tiles in schedule order gives back every layer's weight matrix (fp16-rounded) in the K
    order the kernel's gather uses: eight non-centre taps x channels, [nin_skip,] then the centre tap."""
    from oracle import weights
    import pixelsynth_b200.lmconv as lm

    sd = weights.make_state("lmconv", 0)
    m = lm.LmconvB200(sd, device="cpu")
    dt = np.dtype([("w", np.uint32), ("rows", np.uint16), ("kind", np.uint8), ("t", np.uint8), ("mask", np.uint8),
                   ("cin8", np.uint8), ("kc", np.uint8), ("off", np.uint8), ("col", np.uint16), ("fl", np.uint8), ("gemm", np.uint8)])
    ch = np.frombuffer(m.chunks.numpy().tobytes(), dtype=dt)
    blob = m.wblob.numpy()

    def tile(c):
        rows = int(c["rows"])
        raw = blob[int(c["w"]) * 16:int(c["w"]) * 16 + rows * 128].view(np.float16).reshape(rows, 8, 8)
        src = np.arange(8)[None, :] ^ (np.arange(rows)[:, None] & 7)
        return np.take_along_axis(raw, src[:, :, None], axis=1).reshape(rows, 64).astype(np.float32)

    assert m.plan.n_chunks_total == len(ch) and m.plan.n_chunks_body == len(ch) - 8
    # first GEMM: up_layers.0.u_stream.0.conv_input -- 20 gathered chunks + 3 centre chunks
    w = sd["up_layers.0.u_stream.0.conv_input.weight"].float()
    w9 = w.permute(0, 2, 3, 1).reshape(80, 9, 160)
    nc = torch.cat([w9[:, t] for t in (0, 1, 2, 3, 5, 6, 7, 8)], 1).half().float().numpy()
    got = np.concatenate([tile(ch[i]) for i in range(20)], 1)
    assert np.array_equal(got, nc) and all(ch[i]["kind"] == 0 and ch[i]["kc"] == i for i in range(20))
    ctr = np.concatenate([tile(ch[i]) for i in range(20, 23)], 1)
    assert np.array_equal(ctr[:, :160], w9[:, 4].half().float().numpy()) and not ctr[:, 160:].any()
    assert all(ch[i]["kind"] == 3 for i in range(20, 23)) and ch[22]["fl"] & 2 and m.plan.epi_first[0] == 20
    assert all(ch[i]["gemm"] == 0 for i in range(23)) and ch[23]["gemm"] == 1
    # accumulate flags: only the first chunk of an accumulator overwrites
    assert not ch[0]["fl"] & 1 and all(ch[i]["fl"] & 1 for i in range(1, 23))
    # the issuer waits for the epilogue's operand at the first centre chunk of each of the 32 body GEMMs
    rel = [i for i in range(len(ch)) if ch[i]["fl"] & 16]
    assert rel == [m.plan.epi_first[g] for g in range(32)]
    # nin_out: quarter 0's operand is written into its ring stages, quarters 1-3 re-read it
    assert [int(c["kind"]) for c in ch[-8:]] == [2, 2, 4, 4, 4, 4, 4, 4] and m.plan.raw_mask == sum(1 << t for t in (4, 8, 18, 24))
    # the last quarter of nin_out ends the schedule and completes the logits barrier
    assert ch[-1]["fl"] & 2 and (ch[-1]["fl"] >> 2) & 3 == 2 and ch[-1]["col"] == 384


def test_level_rows_do_not_depend_on_the_worker_count(monkeypatch):
    """ps_lmconv_levels_host splits the images over host threads; rows of a level stay in (image, cell) order."""
    import pixelsynth_b200.lmconv as lmconv

    rng = np.random.default_rng(3)
    B = 24
    m = np.zeros((B, 256, 256), bool)
    yy, xx = np.mgrid[0:256, 0:256]
    for i in range(1, B):   # image 0 has nothing to sample
        m[i] = ((xx - rng.integers(0, 256)) ** 2 + (yy - rng.integers(0, 256)) ** 2) < rng.integers(30, 220) ** 2
    _, order, words, smask = lmconv.glue_host(m)
    outs = []
    for threads in ("1", "2", "5", "6"):
        monkeypatch.setenv("PS_HOST_THREADS", threads)
        for mode, sm in ((0, smask.astype(np.uint8)), (1, None)):
            rows, offs, first_b = lmconv.LmconvB200.levels_host(order, words, sm, mode)
            outs.append((threads, mode, rows.tobytes(), offs.tobytes(), first_b))
    for mode in (0, 1):
        ref = [o for o in outs if o[1] == mode]
        assert all(o[2:] == ref[0][2:] for o in ref), "rows differ between worker counts"
        assert len(ref[0][2]) > 0
