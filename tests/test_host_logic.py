"""Host-side logic: sampling plan / RNG replay vs the oracle port, nearest-index formula vs ATen, configer,
registry."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import contrastiveseg_b200 as cs
from contrastiveseg_b200 import rng
from oracle import ref_port as P
from helpers import load_golden, unpack_perms


def nearest_src(dst, scale, in_size):
    return min(int(np.floor(np.float32(dst) * np.float32(scale))), in_size - 1)


@pytest.mark.parametrize("in_size,out_size", [(512, 128), (1024, 256), (520, 66), (769, 97), (50, 13), (37, 10),
                                              (16, 16), (8, 16), (7, 3)])
def test_nearest_index_formula_matches_aten(in_size, out_size):
    """k_keys uses min(int(floorf(dst * (float)in/out)), in-1); ATen's nearest kernel must agree."""
    src = torch.arange(in_size, dtype=torch.float32).view(1, 1, 1, in_size)
    got = F.interpolate(src, size=(1, out_size), mode="nearest").view(-1).long().tolist()
    scale = np.float32(in_size) / np.float32(out_size)
    mine = [nearest_src(d, scale, in_size) for d in range(out_size)]
    assert mine == got


@pytest.mark.parametrize("name", ["nomem_small", "nomem_oddv", "nomem_mv1", "nomem_nondiv"])
def test_host_plan_and_rank_table_reproduce_reference_sampling(name):
    g = load_golden(name)
    T, bT, ms, mv, K, ign = g["params"].tolist()
    K, ms, mv = int(K), int(ms), int(mv)
    embed = torch.from_numpy(g["embed"])
    B, D, h, w = embed.shape
    lab = P.downsample_labels(torch.from_numpy(g["target"]), h, w).reshape(B, -1)
    prd = torch.from_numpy(g["predict"]).reshape(B, -1)
    counts = np.zeros((B, 2 * K), dtype=np.int64)
    for b in range(B):
        for c in range(K):
            counts[b, 2 * c] = int(((lab[b] == c) & (prd[b] != c)).sum())
            counts[b, 2 * c + 1] = int(((lab[b] == c) & (prd[b] == c)).sum())
    pairs, TC, V = rng.host_plan(counts, ms, mv)
    table = rng.anchor_rank_table(pairs, V, P.PermReplay(unpack_perms(g["perm_flat"], g["perm_lens"])))
    # rebuild the sampled pixel indices from the rank table and compare with the oracle's
    plan = P.sample_anchor_indices(lab, prd, ms, mv, int(ign), P.PermReplay(unpack_perms(g["perm_flat"], g["perm_lens"])))
    idx, cls, img, n_view = plan
    assert (TC, V) == (idx.shape[0], n_view)
    for t, (b, c, nh, ne, kh, ke) in enumerate(pairs):
        hard = ((lab[b] == c) & (prd[b] != c)).nonzero()[:, 0]
        easy = ((lab[b] == c) & (prd[b] == c)).nonzero()[:, 0]
        mine = torch.cat([hard[table[t, :kh].long()], easy[table[t, kh:kh + ke].long()]])
        assert torch.equal(mine, idx[t])
        assert (b, c) == (int(img[t]), int(cls[t]))


def test_configer_protocol():
    c = cs.Configer(cs.cityscapes_contrast_config(with_memory=True))
    assert c.get("contrast", "memory_size") == 5000
    assert c.exists("contrast", "with_memory") and not c.exists("contrast", "nope")
    assert c.get("loss", "params")["ce_ignore_index"] == -1
    with pytest.raises(KeyError):
        c.get("contrast", "nope")


def test_registry_keys_match_reference():
    assert set(cs.SEG_LOSS_DICT) == {"contrast_ce_loss", "contrast_auxce_loss", "mem_contrast_ce_loss"}
    crit = cs.get_seg_loss(cs.Configer(cs.cityscapes_contrast_config()))
    assert isinstance(crit, cs.ContrastCELoss) and crit.loss_weight == 0.1
    assert crit.contrast_criterion.max_samples == 1024 and crit.contrast_criterion.ignore_label == -1
    mem = cs.get_seg_loss(cs.Configer(cs.cityscapes_contrast_config(with_memory=True)))
    assert isinstance(mem, cs.MemContrastCELoss) and mem.with_memory


def test_memory_bank_buffers_match_reference_state_dict():
    bank = cs.MemoryBank(19, 50, 256)
    sd = bank.state_dict()
    assert set(sd) == {"segment_queue", "segment_queue_ptr", "pixel_queue", "pixel_queue_ptr"}
    assert sd["segment_queue"].shape == (19, 50, 256) and sd["segment_queue"].dtype == torch.float32
    assert sd["pixel_queue_ptr"].shape == (19,) and sd["pixel_queue_ptr"].dtype == torch.int64
    assert torch.allclose(sd["pixel_queue"].norm(dim=2), torch.ones(19, 50), atol=1e-5)


def test_bank_rank_table_order():
    counts = np.array([[0, 3, 0, 2], [0, 0, 5, 1]])
    draws = []

    def perm(n):
        draws.append(n)
        return torch.arange(n - 1, -1, -1)
    t = rng.bank_rank_table(counts, 2, perm)
    assert draws == [3, 2, 5, 1]           # image asc, class asc, class 0 skipped, empty slots skipped
    assert t[1].tolist() == [2, 1] and t[3].tolist() == [1, 0] and t[6].tolist() == [4, 3] and t[7].tolist() == [0, 0]


def test_device_rng_model_is_a_bijection_and_statistically_flat():
    """Host model of the device sampler (csrc/pcl_common.cuh keyed_perm): a bijection on [0, n) for every key, and — what
    the multiply/xorshift rounds of the first build failed for small groups — k-subsets whose marginal and PAIR
    frequencies are those of a uniform draw without replacement (chi-square within 5 sigma)."""
    import numpy as np
    from contrastiveseg_b200 import rng
    for n in (1, 2, 3, 7, 60, 100, 1000, 1025):
        for key in (1, 12345678901234567, (1 << 64) - 1):
            assert sorted(rng.keyed_perm(j, n, key) for j in range(n)) == list(range(n))
    for n, k, N in ((60, 10, 3000), (7, 3, 3000), (200, 6, 3000)):
        hits, pairs = np.zeros(n), np.zeros((n, n))
        for s in range(N):
            key = rng.mix64(rng.device_step_seed(123, s + 1) ^ ((1 * 19 + 4) << 1))
            pick = [rng.keyed_perm(j, n, key) for j in range(k)]
            assert len(set(pick)) == k
            hits[pick] += 1
            for a in pick:
                pairs[a, pick] += 1
        exp = N * k / n
        chi = ((hits - exp) ** 2 / exp).sum()
        assert abs(chi - (n - 1)) < 5 * (2 * (n - 1)) ** 0.5, (n, k, chi)
        pe = N * (k / n) * ((k - 1) / (n - 1))
        off = pairs[~np.eye(n, dtype=bool)]
        cells = off.size / 2                                    # symmetric table
        chi_p = ((off - pe) ** 2 / pe).sum() / 2
        assert abs(chi_p - cells) < 6 * (2 * cells) ** 0.5, (n, k, chi_p / cells)


def test_chunked_bank_closed_form_equals_the_oracle_closed_form():
    """tests/helpers.bank_infonce_chunked (the checker of the full-size GPU parity tests) against
    oracle.ref_port.infonce_closed_form on the flattened bank, incl. class-0 anchors (zero-tail positives, Q3) and a
    diagonal inside the class-1 block (Q1)."""
    import torch
    from oracle import ref_port as P
    from helpers import bank_infonce_chunked
    g = torch.Generator().manual_seed(5)
    K, M, D, A = 5, 7, 16, 23
    segq = torch.nn.functional.normalize(torch.randn(K, M, D, generator=g), dim=2)
    pixq = torch.nn.functional.normalize(torch.randn(K, M, D, generator=g), dim=2)
    anchors = torch.nn.functional.normalize(torch.randn(A, D, generator=g), dim=1)
    ya = torch.randint(0, K, (A,), generator=g)
    diag = torch.randperm(A, generator=g)
    contrast, yc = P.flatten_queue(torch.cat((segq, pixq), 1).double())
    ref = P.infonce_closed_form(anchors.double(), ya.double(), contrast, yc, 0.07, 0.07, False, diag_cols=diag)
    got = bank_infonce_chunked(anchors, ya, diag, segq, pixq, 0.07, 0.07, chunk=11)
    assert abs(got["loss"].item() - ref["loss"].item()) <= 1e-12 * abs(ref["loss"].item())
    assert torch.allclose(got["dA"], ref["dA"], rtol=1e-10, atol=1e-14)
