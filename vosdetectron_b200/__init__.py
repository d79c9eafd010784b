"""vosdetectron_b200 -- B200-native (sm_100a) region pipeline of VOSDetectron-style Mask R-CNN.

Hot path only: RPN proposal generation, box NMS, collect/distribute FPN
level assignment, multi-level RoIAlign forward/backward and mask paste-back, as hand-written
CUDA kernels behind a C ABI (include/vosd_b200.h, csrc/), with host-side mirrors of the
reference's Python call signatures:

    modeling.roi_xfrom.roi_align.functions.roi_align.RoIAlignFunction
    modeling.roi_xfrom.roi_align.modules.roi_align.{RoIAlign, RoIAlignAvg, RoIAlignMax}
    modeling.generate_proposals.GenerateProposalsOp
    modeling.generate_anchors.generate_anchors
    modeling.collect_and_distribute_fpn_rpn_proposals.{collect, distribute, CollectAndDistributeFpnRpnProposalsOp}
    modeling.model_builder.roi_feature_transform
    utils.boxes.nms
    core.test.segm_results
    vos_model.flow_align.functions.flow_align.FlowAlignFunction
    vos_model.flow_align.modules.flow_align.FlowAlign

``install_reference_aliases()`` registers them under the reference's own module names so the
unmodified lib/ and lib_vos/ model builders pick them up.
"""
__version__ = "0.1.0"


def install_reference_aliases(legacy_model_roi_align=True):
    """Make ``import modeling.roi_xfrom.roi_align.functions.roi_align`` (as done at
    lib/modeling/model_builder.py:13, lib_vos/vos_modeling/vos_model_builder.py:13 and
    generalized_rcnn_predictor_with_boxes.py:14) resolve to this package.  Call it before the
    reference's model builders are imported; reference packages already imported are patched in
    place (attribute replacement) instead."""
    import importlib
    import sys

    from .modeling.roi_xfrom.roi_align.functions import roi_align as fn_mod
    from .modeling.roi_xfrom.roi_align.modules import roi_align as mod_mod

    def put(name, module):
        sys.modules[name] = module
        parent, _, leaf = name.rpartition('.')
        if parent in sys.modules:
            setattr(sys.modules[parent], leaf, module)

    put('modeling.roi_xfrom.roi_align.functions.roi_align', fn_mod)
    put('modeling.roi_xfrom.roi_align.modules.roi_align', mod_mod)
    # lib_vos/vos_modeling/vos_model_builder.py:22 imports the module class; FAN.py:19 the package
    from .vos_model import flow_align as fa_pkg
    from .vos_model.flow_align.functions import flow_align as fa_fn
    from .vos_model.flow_align.modules import flow_align as fa_mod
    put('vos_model.flow_align', fa_pkg)
    put('vos_model.flow_align.functions.flow_align', fa_fn)
    # the cffi-level modules the reference's Function classes import (`from .._ext import roi_align / flow_align`)
    from .modeling.roi_xfrom.roi_align._ext import roi_align as ra_ext
    from .vos_model.flow_align._ext import flow_align as fa_ext
    put('modeling.roi_xfrom.roi_align._ext.roi_align', ra_ext)
    put('vos_model.flow_align._ext.flow_align', fa_ext)
    put('vos_model.flow_align.modules.flow_align', fa_mod)
    if legacy_model_roi_align:
        # lib/model/roi_align is the dead 3-argument variant with different maths
        # (src/roi_align_kernel.cu:40-46); alias only the 4-argument form.
        put('model.roi_align.functions.roi_align', fn_mod)
    patches = {
        'modeling.generate_proposals': ('.modeling.generate_proposals', ['GenerateProposalsOp']),
        'modeling.collect_and_distribute_fpn_rpn_proposals':
            ('.modeling.collect_and_distribute_fpn_rpn_proposals',
             ['collect', 'distribute', 'CollectAndDistributeFpnRpnProposalsOp']),
        'utils.boxes': ('.utils.boxes', ['nms', 'bbox_overlaps']),
        'core.test': ('.core.test', ['segm_results']),
        # lib_vos/tools/vos_test.py is imported as the top-level module `vos_test` (infer_davis_sequential.py:27)
        'vos_test': ('.core.vos_test', ['segm_results', 'box_results_with_nms_and_limit', 'nms_with_mask_iou']),
    }
    for ref_name, (mine, names) in patches.items():
        m = importlib.import_module(mine, __name__)
        if ref_name in sys.modules:
            for n in names:
                setattr(sys.modules[ref_name], n, getattr(m, n))
    return sorted(k for k in sys.modules if k.startswith(('modeling.roi_xfrom', 'model.roi_align', 'vos_model.flow_align')))
