"""torch-CPU restatement of spconv v1.2's native convolution algorithm (`ConvAlgo.Native`: per kernel offset
gather rows -> SGEMM -> scatter-add; upstream src/spconv/spconv_ops.cc indiceConv / indiceConvBackward, not vendored
in /root/reference -- spec: SURVEY.md Appendix A.5).  Call sites in the reference: model/unet.py:36,
model/unet_block.py:20,26,29,48,70,78."""
import numpy as np
import torch


def indice_conv_ref(features, filters, pairs, pairnum, n_out, inverse=False, subm=False):
    """features [M_in,Cin]; filters [k,k,k,Cin,Cout]; pairs [2,K,M]; -> [n_out,Cout].
    Differentiable through torch autograd (index_select / mm / index_add)."""
    Cin, Cout = filters.shape[-2], filters.shape[-1]
    W = filters.reshape(-1, Cin, Cout)
    K = W.shape[0]
    pairs = torch.as_tensor(np.asarray(pairs), dtype=torch.int64)
    pairnum = [int(v) for v in np.asarray(pairnum)]
    out = torch.zeros((n_out, Cout), dtype=features.dtype)
    kstar = -1
    if subm:  # centre offset handled densely first: out = feat @ W[k*], k* = argmax(pairnum)
        kstar = int(np.argmax(pairnum))
        out = out + features @ W[kstar]
    gi, si = (1, 0) if inverse else (0, 1)
    for k in range(K):
        n = pairnum[k]
        if n == 0 or k == kstar:
            continue
        buf = features.index_select(0, pairs[gi, k, :n])
        out = out.index_add(0, pairs[si, k, :n], buf @ W[k])
    return out


def indice_conv_backward_ref(features, filters, out_bp, pairs, pairnum, inverse=False, subm=False):
    """-> (input_bp [M_in,Cin], filters_bp like filters).  dW[k] = buf^T @ dout[pairs_out];
    din[pairs_in] += dout[pairs_out] @ W[k]^T; SubM centre handled densely (A.5)."""
    Cin, Cout = filters.shape[-2], filters.shape[-1]
    W = filters.reshape(-1, Cin, Cout)
    K = W.shape[0]
    pairs = torch.as_tensor(np.asarray(pairs), dtype=torch.int64)
    pairnum = [int(v) for v in np.asarray(pairnum)]
    din = torch.zeros_like(features)
    dW = torch.zeros_like(W)
    kstar = -1
    if subm:
        kstar = int(np.argmax(pairnum))
        dW[kstar] = features.t() @ out_bp
        din = din + out_bp @ W[kstar].t()
    gi, si = (1, 0) if inverse else (0, 1)
    for k in range(K):
        n = pairnum[k]
        if n == 0 or k == kstar:
            continue
        a = features.index_select(0, pairs[gi, k, :n])
        g = out_bp.index_select(0, pairs[si, k, :n])
        dW[k] = a.t() @ g
        din = din.index_add(0, pairs[gi, k, :n], g @ W[k].t())
    return din, dW.reshape(filters.shape)
