"""User-level point<->voxel operators with the names and argument meaning of the reference's
lib/pointgroup_ops/functions/pointgroup_ops.py (Voxelization_Idx 13-41, Voxelization 44-77, PointRecover 80-114,
BallQueryBatchP 117-152, SecMean/SecMin/SecMax 258-347, ...), written against doda_b200.pg_op.

The reference's own wrapper file also works unchanged on top of compat/PG_OP.py; this module exists so that the
engine can be used (bench, tests, smoke) on a box where /root/reference is absent.
"""
import torch
from torch.autograd import Function

from . import pg_op as PG_OP
from ._lib import lib, check


def voxelization_idx(coords, batchsize, mode=4):
    """coords int64 [N, 3|4] on CPU -> (output_coords int64 [M,ncol], input_map int32 [N] (p2v),
    output_map int32 [M, 1+maxActive] (v2p))."""
    assert coords.is_contiguous()
    N = coords.size(0)
    output_coords = coords.new_empty(0)
    input_map = torch.zeros(N, dtype=torch.int32)
    output_map = torch.zeros(0, dtype=torch.int32)
    PG_OP.voxelize_idx(coords, output_coords, input_map, output_map, batchsize, mode)
    return output_coords, input_map, output_map


_vox_pinned = [None, 0]


def voxelization_idx_gpu(coords, batchsize, mode=4):
    """`voxelization_idx` on the device: coords int64 [N, 3 or 4] CUDA ->
    (output_coords int64 [M, ncol], input_index int32 [N], output_index int32 [M, 1 + maxActive]), all CUDA and
    bit-identical to the CPU / reference result (first-touch voxel order, ascending points per voxel).
    One host sync (M and maxActive size the outputs); run it on a side stream a batch ahead to hide it."""
    from . import ops as _ops
    if not coords.is_cuda:
        raise RuntimeError("voxelization_idx_gpu needs a CUDA tensor (the CPU entry point is voxelization_idx)")
    assert coords.dtype == torch.int64 and coords.dim() == 2 and coords.is_contiguous()
    N, ncol = coords.shape
    dev = coords.device
    input_map = torch.empty(N, dtype=torch.int32, device=dev)
    ws = _ops._workspace(int(lib.b200sp_voxelize_idx_gpu_ws_bytes(N)), dev, "vox")
    if _vox_pinned[0] is None:
        _vox_pinned[0] = torch.zeros(64, 4, dtype=torch.int32).pin_memory()
    _vox_pinned[1] = (_vox_pinned[1] + 1) % 64
    info = _vox_pinned[0][_vox_pinned[1]]
    st = _ops._stream()
    check(lib.b200sp_voxelize_idx_gpu_begin(coords.data_ptr(), N, ncol, int(mode), input_map.data_ptr(), info.data_ptr(),
                                            ws.data_ptr(), ws.numel(), st), "voxelize_idx_gpu_begin")
    torch.cuda.current_stream(dev).synchronize()
    M, A, bad, dup = [int(v) for v in info.tolist()]
    if bad:
        raise RuntimeError("voxelization_idx_gpu: coordinates must be non-negative, x, y, z < 2^20, and the batch index must fit "
                           "into the 64 - 3 * bits(max coordinate) key bits the grid leaves (any batch for grids up to 2^17 wide)")
    if int(mode) == 0 and dup:
        raise RuntimeError("libb200sparse voxelize_idx failed: mode 0 requires unique coordinates")
    A = A if int(mode) in (3, 4) else 1
    out_coords = torch.empty((M, ncol), dtype=torch.int64, device=dev)
    output_map = torch.empty((M, A + 1), dtype=torch.int32, device=dev)
    check(lib.b200sp_voxelize_idx_gpu_finish(coords.data_ptr(), N, ncol, int(mode), M, A, out_coords.data_ptr(),
                                             output_map.data_ptr(), ws.data_ptr(), ws.numel(), st),
          "voxelize_idx_gpu_finish")
    return out_coords, input_map, output_map


class Voxelization(Function):
    @staticmethod
    def forward(ctx, feats, map_rule, mode=4):
        assert map_rule.is_contiguous() and feats.is_contiguous()
        N, C = feats.size()
        M = map_rule.size(0)
        maxActive = map_rule.size(1) - 1
        output_feats = torch.zeros((M, C), dtype=torch.float32, device=feats.device)
        ctx.for_backwards = (map_rule, mode, maxActive, N)
        PG_OP.voxelize_fp(feats, output_feats, map_rule, mode, M, maxActive, C)
        return output_feats

    @staticmethod
    def backward(ctx, d_output_feats):
        map_rule, mode, maxActive, N = ctx.for_backwards
        M, C = d_output_feats.size()
        d_feats = torch.zeros((N, C), dtype=torch.float32, device=d_output_feats.device)
        PG_OP.voxelize_bp(d_output_feats.contiguous(), d_feats, map_rule, mode, M, maxActive, C)
        return d_feats, None, None


voxelization = Voxelization.apply


class PointRecover(Function):
    @staticmethod
    def forward(ctx, feats, map_rule, nPoint):
        assert map_rule.is_contiguous() and feats.is_contiguous()
        M, C = feats.size()
        maxActive = map_rule.size(1) - 1
        output_feats = torch.zeros((nPoint, C), dtype=torch.float32, device=feats.device)
        ctx.for_backwards = (map_rule, maxActive, M)
        PG_OP.point_recover_fp(feats, output_feats, map_rule, M, maxActive, C)
        return output_feats

    @staticmethod
    def backward(ctx, d_output_feats):
        map_rule, maxActive, M = ctx.for_backwards
        N, C = d_output_feats.size()
        d_feats = torch.zeros((M, C), dtype=torch.float32, device=d_output_feats.device)
        PG_OP.point_recover_bp(d_output_feats.contiguous(), d_feats, map_rule, M, maxActive, C)
        return d_feats, None, None


point_recover = PointRecover.apply


def ballquery_batch_p(coords, batch_idxs, batch_offsets, radius, meanActive):
    """-> (idx int32 [nActive], start_len int32 [n,2]); retries with a larger buffer like the reference (137-143)."""
    n = coords.size(0)
    assert coords.is_contiguous() and coords.is_cuda
    while True:
        idx = torch.zeros(n * meanActive, dtype=torch.int32, device=coords.device)
        start_len = torch.zeros((n, 2), dtype=torch.int32, device=coords.device)
        nActive = PG_OP.ballquery_batch_p(coords, batch_idxs, batch_offsets, idx, start_len, n, meanActive, radius)
        if nActive <= n * meanActive:
            break
        meanActive = int(nActive // n + 1)
    return idx[:nActive], start_len


def bfs_cluster(semantic_label, ball_query_idxs, start_len, threshold):
    N = start_len.size(0)
    cluster_idxs = semantic_label.new_empty(0)
    cluster_offsets = semantic_label.new_empty(0)
    PG_OP.bfs_cluster(semantic_label, ball_query_idxs, start_len, cluster_idxs, cluster_offsets, N, threshold)
    return cluster_idxs, cluster_offsets


class SecMean(Function):
    @staticmethod
    def forward(ctx, inp, offsets):
        nProposal = offsets.size(0) - 1
        C = inp.size(1)
        out = torch.zeros((nProposal, C), dtype=torch.float32, device=inp.device)
        PG_OP.sec_mean(inp.contiguous(), offsets.contiguous(), out, nProposal, C)
        ctx.for_backwards = (offsets, inp.size(0))
        return out

    @staticmethod
    def backward(ctx, d_out):
        offsets, N = ctx.for_backwards
        nProposal, C = d_out.size()
        d_inp = torch.zeros((N, C), dtype=torch.float32, device=d_out.device)
        PG_OP.sec_mean_bp(d_inp, offsets, d_out.contiguous(), nProposal, C)
        return d_inp, None


sec_mean = SecMean.apply


def sec_min(inp, offsets):
    nProposal, C = offsets.size(0) - 1, inp.size(1)
    out = torch.zeros((nProposal, C), dtype=torch.float32, device=inp.device)
    PG_OP.sec_min(inp.contiguous(), offsets.contiguous(), out, nProposal, C)
    return out


def sec_max(inp, offsets):
    nProposal, C = offsets.size(0) - 1, inp.size(1)
    out = torch.zeros((nProposal, C), dtype=torch.float32, device=inp.device)
    PG_OP.sec_max(inp.contiguous(), offsets.contiguous(), out, nProposal, C)
    return out


class RoiPool(Function):
    @staticmethod
    def forward(ctx, feats, proposals_offset):
        nProposal = proposals_offset.size(0) - 1
        sumNPoint, C = feats.size()
        out = torch.zeros((nProposal, C), dtype=torch.float32, device=feats.device)
        maxidx = torch.zeros((nProposal, C), dtype=torch.int32, device=feats.device)
        PG_OP.roipool_fp(feats.contiguous(), proposals_offset.contiguous(), out, maxidx, nProposal, C)
        ctx.for_backwards = (maxidx, proposals_offset, sumNPoint)
        return out

    @staticmethod
    def backward(ctx, d_output_feats):
        nProposal, C = d_output_feats.size()
        maxidx, proposals_offset, sumNPoint = ctx.for_backwards
        d_feats = torch.zeros((sumNPoint, C), dtype=torch.float32, device=d_output_feats.device)
        PG_OP.roipool_bp(d_feats, proposals_offset, maxidx, d_output_feats.contiguous(), nProposal, C)
        return d_feats, None


roipool = RoiPool.apply


def get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum):
    nInstance = instance_pointnum.size(0)
    nProposal = proposals_offset.size(0) - 1
    iou = torch.zeros((nProposal, nInstance), dtype=torch.float32, device=proposals_idx.device)
    PG_OP.get_iou(proposals_idx, proposals_offset, instance_labels, instance_pointnum, iou, nInstance, nProposal)
    return iou


def knn_batch(xyz, query_xyz, batch_idxs, query_batch_offsets, k):
    n, m = xyz.size(0), query_xyz.size(0)
    idx = torch.zeros((n, k), dtype=torch.int32, device=xyz.device)
    PG_OP.knn_batch(xyz, query_xyz, batch_idxs, query_batch_offsets, idx, n, m, k)
    return idx
