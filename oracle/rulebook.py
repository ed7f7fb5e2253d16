"""numpy restatement of spconv v1.2 `get_indice_pairs` (upstream src/spconv/indice.cc getIndicePairsConv /
getIndicePairsSubM with getValidOutPos; not vendored in /root/reference -- see SURVEY.md Appendix A.2-A.4, which
is the spec this file follows).  Output is in the canonical order of A.4: output sites in ascending flattened
index (SubM: the input order), pairs inside each offset ascending in input row."""
import numpy as np


def _triple(v):
    return [int(x) for x in v] if isinstance(v, (list, tuple, np.ndarray)) else [int(v)] * 3


def conv_out_shape(shape, ks, st, pd, dl):
    # A.2: O = (S + 2p - d(k-1) - 1)//s + 1
    return [(s + 2 * p - d * (k - 1) - 1) // t + 1 for s, k, t, p, d in zip(shape, ks, st, pd, dl)]


def get_indice_pairs_ref(indices, batch_size, spatial_shape, ksize, stride=1, padding=0, dilation=1, subm=False):
    """indices int [M,4] (batch,i0,i1,i2) -> (outids int32 [M',4], pairs int32 [2,K,M] (-1 padded),
    pairnum int32 [K], out_shape)."""
    idx = np.asarray(indices, dtype=np.int64)
    M = idx.shape[0]
    ks, st, pd, dl = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    S = [int(s) for s in spatial_shape]
    if subm:
        st = [1, 1, 1]
        pd = [k // 2 for k in ks]  # spconv forces padding = k//2, stride = 1 for SubM
        O = S
    else:
        O = conv_out_shape(S, ks, st, pd, dl)
    K = ks[0] * ks[1] * ks[2]
    b, x = idx[:, 0], idx[:, 1:4]
    cand_key = np.full((K, M), -1, dtype=np.int64)
    for k in range(K):
        kap = (k // (ks[1] * ks[2]), (k // ks[2]) % ks[1], k % ks[2])  # A.3: last axis fastest
        ok = np.ones(M, dtype=bool)
        o = np.zeros((M, 3), dtype=np.int64)
        for a in range(3):
            num = x[:, a] + pd[a] - kap[a] * dl[a]  # kappa = (x - o*s + p)/d  <=>  o*s = x + p - kappa*d
            ok &= (num >= 0) & (num % st[a] == 0)
            o[:, a] = num // st[a]
            ok &= o[:, a] <= O[a] - 1
        key = ((b * O[0] + o[:, 0]) * O[1] + o[:, 1]) * O[2] + o[:, 2]
        cand_key[k] = np.where(ok, key, -1)
    if subm:
        in_key = ((b * S[0] + x[:, 0]) * S[1] + x[:, 1]) * S[2] + x[:, 2]
        order = np.argsort(in_key, kind="stable")
        skeys = in_key[order]
        outids = idx.astype(np.int32)
        out_keys_sorted, out_rows_sorted = skeys, order
    else:
        uniq = np.unique(cand_key[cand_key >= 0])  # ascending flattened index (spconv CUDA path: torch::_unique)
        out_keys_sorted, out_rows_sorted = uniq, np.arange(uniq.shape[0])
        oc = np.zeros((uniq.shape[0], 4), dtype=np.int64)
        t = uniq.copy()
        oc[:, 3] = t % O[2]; t //= O[2]
        oc[:, 2] = t % O[1]; t //= O[1]
        oc[:, 1] = t % O[0]; t //= O[0]
        oc[:, 0] = t
        outids = oc.astype(np.int32)
    pairs = np.full((2, K, M), -1, dtype=np.int32)
    pairnum = np.zeros(K, dtype=np.int32)
    for k in range(K):
        key = cand_key[k]
        pos = np.searchsorted(out_keys_sorted, key)
        pos = np.clip(pos, 0, max(out_keys_sorted.shape[0] - 1, 0))
        hit = (key >= 0) & (out_keys_sorted.shape[0] > 0)
        if out_keys_sorted.shape[0] > 0:
            hit &= out_keys_sorted[pos] == key
        j = np.nonzero(hit)[0]  # ascending input row
        n = j.shape[0]
        pairnum[k] = n
        pairs[0, k, :n] = j
        pairs[1, k, :n] = out_rows_sorted[pos[j]]
    return outids, pairs, pairnum, O


def canonicalize(outids, pairs, pairnum, out_shape, subm=False):
    """Bring any valid spconv rulebook (arbitrary pair order / output numbering) to the canonical form of A.4."""
    outids = np.asarray(outids, dtype=np.int64)
    pairs = np.asarray(pairs).copy()
    K, M = pairs.shape[1], pairs.shape[2]
    if subm:
        perm = np.arange(outids.shape[0])
        new_out = outids
    else:
        O = out_shape
        key = ((outids[:, 0] * O[0] + outids[:, 1]) * O[1] + outids[:, 2]) * O[2] + outids[:, 3]
        order = np.argsort(key, kind="stable")
        perm = np.empty_like(order)
        perm[order] = np.arange(order.shape[0])
        new_out = outids[order]
    res = np.full_like(pairs, -1)
    for k in range(K):
        n = int(pairnum[k])
        pi, po = pairs[0, k, :n], perm[pairs[1, k, :n]] if n else pairs[1, k, :0]
        o = np.argsort(pi, kind="stable")
        res[0, k, :n], res[1, k, :n] = pi[o], po[o]
    return new_out.astype(np.int32), res, np.asarray(pairnum).astype(np.int32)
