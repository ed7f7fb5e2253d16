"""ReLU-gate capture for the pinned-gate comparison -- TEST INFRASTRUCTURE (used by tests/ and __graft_entry__.smoke()).

Whole-net gradients of a 71-conv / 65-BN ReLU net are discontinuous in the rounding of the forward pass: any two fp32
evaluations flip a ~1e-6 fraction of the gates and each flip changes its gradient contribution by 100 % (DESIGN.md
section 4).  The comparison that isolates the arithmetic is therefore: record which BN+ReLU outputs the engine found
positive, and evaluate the fp64 oracle with exactly those gates (oracle/unet_ref.py: model_step_ref(relu_masks=...)).
"""


def capture_relu_masks(model):
    """Instrument a SparseConvNet (the mirror or the reference's own class) built on the engine's spconv surface so
    that its next forward records, per BatchNorm key (state_dict prefix), which BN+ReLU outputs were > 0.
    -> dict filled during the forward.  Test infrastructure: wraps the fused-triplet entry point per conv instance."""
    from doda_b200 import spconv
    from doda_b200.spconv.modules import is_sparse_conv, _is_bn_like
    masks = {}
    # a U-Net sub-tree that runs as one taped autograd node (doda_b200/tape.py) records its gates itself
    from doda_b200 import tape as _tape
    _tape.capture = (masks, {id(m): name for name, m in model.named_modules() if _is_bn_like(m)})
    for name, seq in model.named_modules():
        if not isinstance(seq, spconv.SparseSequential):
            continue
        mods = list(seq._modules.items())
        for i, (k, m) in enumerate(mods):
            if not _is_bn_like(m):
                continue
            key = "%s.%s" % (name, k)
            if i + 2 < len(mods) and is_sparse_conv(mods[i + 2][1]):
                conv = mods[i + 2][1]

                def wrapped(input, bn, stats_args, _orig=conv.forward_after_bn_relu, _key=key):
                    out = _orig(input, bn, stats_args)
                    masks[_key] = (input.features.detach() > 0).cpu()
                    return out
                conv.forward_after_bn_relu = wrapped
            else:  # BN + ReLU at the end of a sequential (output_layer): the sequential's output holds the activation
                seq.register_forward_hook(lambda mod, inp, out, _key=key: masks.__setitem__(_key, (out.features.detach() > 0).cpu()))
    return masks
