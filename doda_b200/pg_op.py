"""Module with the exact function surface of the reference's `PG_OP` torch extension
(lib/pointgroup_ops/src/pointgroup_ops_api.cpp:7-26), backed by libb200sparse.so.

`compat/PG_OP.py` aliases this module so that the reference's unchanged wrapper file
lib/pointgroup_ops/functions/pointgroup_ops.py (`import PG_OP`, line 11) binds to it.
Ownership follows the reference: the caller allocates (zeroed) outputs, except voxelize_idx and bfs_cluster
which resize the empty tensors they are given (voxelize.cpp:22-26, bfs_cluster.cpp:103-106).
"""
import ctypes

import torch

from ._lib import lib, check


def _s():
    from .ops import _stream
    return _stream()


def _cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("PG_OP: expected a CUDA tensor (no CPU fallback for the GPU entry points)")
        if not t.is_contiguous():
            raise RuntimeError("PG_OP: tensors must be contiguous")


def voxelize_idx(coords, output_coords, input_map, output_map, batchSize, mode):
    """CPU. coords int64 [N,3|4]; resizes output_coords [M,ncol] int64 and output_map [M,1+maxActive] int32;
    fills input_map int32 [N]."""
    if coords.is_cuda:
        raise RuntimeError("voxelize_idx is the CPU (DataLoader-side) entry point")
    assert coords.dim() == 2 and coords.dtype == torch.int64 and coords.is_contiguous()
    assert input_map.dtype == torch.int32 and input_map.numel() == coords.shape[0]
    N, ncol = coords.shape
    M = ctypes.c_int64(0)
    A = ctypes.c_int32(0)
    # one hashing pass: the size query also writes the point -> voxel map, _fill builds the rest from it
    check(lib.b200sp_voxelize_idx_cpu(coords.data_ptr(), N, ncol, int(batchSize), int(mode), None,
                                      input_map.data_ptr(), None, ctypes.byref(M), ctypes.byref(A)), "voxelize_idx(size)")
    output_coords.resize_(M.value, ncol)
    output_map.resize_(M.value, A.value + 1)
    check(lib.b200sp_voxelize_idx_cpu_fill(coords.data_ptr(), N, ncol, int(mode), input_map.data_ptr(), M.value, A.value,
                                           output_coords.data_ptr(), output_map.data_ptr()), "voxelize_idx(fill)")


def voxelize_fp(feats, output_feats, output_map, mode, nActive, maxActive, nPlane):
    _cuda(feats, output_feats, output_map)
    check(lib.b200sp_voxelize_fp(feats.data_ptr(), output_feats.data_ptr(), output_map.data_ptr(),
                                 1 if mode == 4 else 0, nActive, maxActive, nPlane, _s()), "voxelize_fp")


def voxelize_bp(d_output_feats, d_feats, output_map, mode, nActive, maxActive, nPlane):
    _cuda(d_output_feats, d_feats, output_map)
    check(lib.b200sp_voxelize_bp(d_output_feats.data_ptr(), d_feats.data_ptr(), output_map.data_ptr(),
                                 1 if mode == 4 else 0, nActive, maxActive, nPlane, _s()), "voxelize_bp")


def point_recover_fp(feats, output_feats, idx_map, nActive, maxActive, nPlane):
    # voxel -> point broadcast == the voxelize_bp kernel without averaging (voxelize.cpp:185-193)
    _cuda(feats, output_feats, idx_map)
    check(lib.b200sp_voxelize_bp(feats.data_ptr(), output_feats.data_ptr(), idx_map.data_ptr(), 0, nActive,
                                 maxActive, nPlane, _s()), "point_recover_fp")


def point_recover_bp(d_output_feats, d_feats, idx_map, nActive, maxActive, nPlane):
    _cuda(d_output_feats, d_feats, idx_map)
    check(lib.b200sp_voxelize_fp(d_output_feats.data_ptr(), d_feats.data_ptr(), idx_map.data_ptr(), 0, nActive,
                                 maxActive, nPlane, _s()), "point_recover_bp")


def ballquery_batch_p(xyz, batch_idxs, batch_offsets, idx, start_len, n, meanActive, radius):
    _cuda(xyz, batch_idxs, batch_offsets, idx, start_len)
    wsb = lib.b200sp_ballquery_ws_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=xyz.device)
    total = ctypes.c_int32(0)
    check(lib.b200sp_ballquery_batch_p(xyz.data_ptr(), batch_idxs.data_ptr(), batch_offsets.data_ptr(),
                                       idx.data_ptr(), start_len.data_ptr(), n, meanActive, float(radius),
                                       ws.data_ptr(), wsb, ctypes.byref(total), _s()), "ballquery_batch_p")
    return int(total.value)


def bfs_cluster(semantic_label, ball_query_idxs, start_len, cluster_idxs, cluster_offsets, N, threshold):
    for t in (semantic_label, ball_query_idxs, start_len):
        if t.is_cuda:
            raise RuntimeError("bfs_cluster is a CPU entry point (bfs_cluster.cpp:93)")
        assert t.dtype == torch.int32 and t.is_contiguous()
    ni, nc = ctypes.c_int64(0), ctypes.c_int64(0)
    check(lib.b200sp_bfs_cluster_cpu(semantic_label.data_ptr(), ball_query_idxs.data_ptr(), start_len.data_ptr(),
                                     N, threshold, None, None, ctypes.byref(ni), ctypes.byref(nc)), "bfs(size)")
    cluster_idxs.resize_(ni.value, 2)
    cluster_offsets.resize_(nc.value + 1)
    cluster_idxs.zero_()
    cluster_offsets.zero_()
    check(lib.b200sp_bfs_cluster_cpu(semantic_label.data_ptr(), ball_query_idxs.data_ptr(), start_len.data_ptr(),
                                     N, threshold, cluster_idxs.data_ptr(), cluster_offsets.data_ptr(),
                                     ctypes.byref(ni), ctypes.byref(nc)), "bfs_cluster")


def roipool_fp(feats, proposals_offset, output_feats, output_maxidx, nProposal, C):
    _cuda(feats, proposals_offset, output_feats, output_maxidx)
    check(lib.b200sp_roipool_fp(feats.data_ptr(), proposals_offset.data_ptr(), output_feats.data_ptr(),
                                output_maxidx.data_ptr(), nProposal, C, _s()), "roipool_fp")


def roipool_bp(d_feats, proposals_offset, output_maxidx, d_output_feats, nProposal, C):
    _cuda(d_feats, output_maxidx, d_output_feats)
    check(lib.b200sp_roipool_bp(d_output_feats.data_ptr(), output_maxidx.data_ptr(), d_feats.data_ptr(), nProposal,
                                C, _s()), "roipool_bp")


def get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou, nInstance, nProposal):
    _cuda(proposals_idx, proposals_offset, instance_labels, instance_pointnum, proposals_iou)
    assert instance_labels.dtype == torch.int64
    check(lib.b200sp_get_iou(proposals_idx.data_ptr(), proposals_offset.data_ptr(), instance_labels.data_ptr(),
                             instance_pointnum.data_ptr(), proposals_iou.data_ptr(), nProposal, nInstance, _s()),
          "get_iou")


def sec_mean(inp, offsets, out, nProposal, C):
    _cuda(inp, offsets, out)
    check(lib.b200sp_sec_mean(inp.data_ptr(), offsets.data_ptr(), out.data_ptr(), nProposal, C, _s()), "sec_mean")


def sec_mean_bp(d_inp, offsets, d_out, nProposal, C):
    _cuda(d_inp, offsets, d_out)
    check(lib.b200sp_sec_mean_bp(d_out.data_ptr(), offsets.data_ptr(), d_inp.data_ptr(), nProposal, C, _s()),
          "sec_mean_bp")


def sec_min(inp, offsets, out, nProposal, C):
    _cuda(inp, offsets, out)
    check(lib.b200sp_sec_min(inp.data_ptr(), offsets.data_ptr(), out.data_ptr(), nProposal, C, _s()), "sec_min")


def sec_max(inp, offsets, out, nProposal, C):
    _cuda(inp, offsets, out)
    check(lib.b200sp_sec_max(inp.data_ptr(), offsets.data_ptr(), out.data_ptr(), nProposal, C, _s()), "sec_max")


def knn_batch(xyz, query_xyz, batch_idxs, query_batch_offsets, idx, n, m, k):
    _cuda(xyz, query_xyz, batch_idxs, query_batch_offsets, idx)
    check(lib.b200sp_knn_batch(xyz.data_ptr(), query_xyz.data_ptr(), batch_idxs.data_ptr(),
                               query_batch_offsets.data_ptr(), idx.data_ptr(), n, m, k, _s()), "knn_batch")
