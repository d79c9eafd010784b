"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Runs the *unmodified* reference hot path (YeLyuUT/VOSDetectron, mounted read-only
at /root/reference) inside this container so that

  * ``tests/golden/make_golden.py`` can dump golden input/output vectors, and
  * ``tests/test_oracle_vs_reference.py`` can pin the restated oracle
    (``oracle/region_oracle.py`` + ``oracle/oracle.c``) against the real code.

/root/reference does not exist on the GPU box, so nothing here may be reached
from ``-m gpu`` tests, ``smoke()`` or ``bench.py``.  ``available()`` says whether
the reference tree is present.

The shim list follows SURVEY.md Appendix A (nothing under /root/reference is
modified):
  * ``np.int`` / ``np.float`` aliases (removed in NumPy >= 1.24; used at
    lib/modeling/generate_anchors.py:63-72 and lib/core/test.py:904-905),
  * ``yaml.load`` defaulting to SafeLoader (lib/core/config.py:1110),
  * stub modules ``nn``, ``matplotlib``, ``pycocotools*``,
    ``datasets.json_dataset``, ``imdb.vos.davis_db`` (imported at module scope by
    lib/core/config.py:19, lib/modeling/collect_and_distribute_fpn_rpn_proposals.py:4,16,
    lib/core/test.py),
  * ``utils.cython_nms`` / ``utils.cython_bbox`` rebuilt from the reference
    ``.pyx`` sources by ``oracle/build_ref.py`` into ``oracle/_ref/``.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("VOSD_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD = os.path.join(HERE, "_ref")

_state = {"ready": False}


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "modeling"))


def _load_ext(name, fname_prefix):
    for f in sorted(os.listdir(REF_BUILD)):
        if f.startswith(fname_prefix) and f.endswith(".so"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF_BUILD, f))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    raise ImportError("%s*.so not found in %s (run python oracle/build_ref.py)" % (fname_prefix, REF_BUILD))


def load_cython_nms():
    """The reference's cython_nms, compiled from its own .pyx (oracle/_ref)."""
    return _load_ext("cython_nms", "cython_nms")


def setup():
    """Make ``import modeling.generate_proposals`` etc. work; idempotent."""
    if _state["ready"]:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    import yaml
    if not getattr(yaml.load, "_vosd_wrapped", False):
        _orig = yaml.load

        def _load(stream, Loader=None):
            return _orig(stream, Loader=Loader or yaml.SafeLoader)
        _load._vosd_wrapped = True
        yaml.load = _load

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    stub("nn")
    stub("matplotlib", use=lambda *a, **k: None)
    stub("matplotlib.pyplot")
    stub("pycocotools")
    sys.modules["pycocotools"].mask = stub("pycocotools.mask", encode=None)
    stub("pycocotools.coco", COCO=object)
    stub("pycocotools.cocoeval", COCOeval=object)

    sys.path.insert(0, os.path.join(REF_ROOT, "lib"))
    import utils  # noqa: the reference's lib/utils package
    nms_mod = _load_ext("utils.cython_nms", "cython_nms")
    bbox_mod = _load_ext("utils.cython_bbox", "cython_bbox")
    sys.modules["utils.cython_nms"] = nms_mod
    sys.modules["utils.cython_bbox"] = bbox_mod
    utils.cython_nms = nms_mod
    utils.cython_bbox = bbox_mod
    import datasets  # noqa
    stub("datasets.json_dataset")
    datasets.json_dataset = sys.modules["datasets.json_dataset"]
    stub("imdb")
    stub("imdb.vos")
    stub("imdb.vos.davis_db")
    sys.modules["imdb"].vos = sys.modules["imdb.vos"]
    sys.modules["imdb.vos"].davis_db = sys.modules["imdb.vos.davis_db"]
    _state["ready"] = True


def ref():
    """Namespace with the reference callables on the hot path."""
    setup()
    ns = types.SimpleNamespace()
    from core.config import cfg, merge_cfg_from_file
    import utils.boxes as box_utils
    import utils.fpn as fpn_utils
    from modeling.generate_anchors import generate_anchors
    from modeling.generate_proposals import GenerateProposalsOp
    ns.cfg = cfg
    ns.merge_cfg_from_file = merge_cfg_from_file
    ns.box_utils = box_utils
    ns.fpn_utils = fpn_utils
    ns.generate_anchors = generate_anchors
    ns.GenerateProposalsOp = GenerateProposalsOp
    from modeling.collect_and_distribute_fpn_rpn_proposals import collect, distribute
    ns.collect = collect
    ns.distribute = distribute
    import core.test as core_test
    ns.core_test = core_test
    ns.yaml_r50 = os.path.join(REF_ROOT, "configs/baselines/e2e_mask_rcnn_R-50-FPN_1x.yaml")
    return ns


def segm_results_capture(r, cls_boxes, masks, ref_boxes, im_h, im_w):
    """Run the reference segm_results (lib/core/test.py:801-855) unmodified and
    capture every full-frame uint8 mask it hands to pycocotools' encode."""
    captured = []

    def enc(a):
        captured.append(np.array(a[:, :, 0]))
        return [{"counts": b"x"}]
    r.core_test.mask_util.encode = enc
    r.core_test.segm_results(cls_boxes, masks, ref_boxes, im_h, im_w)
    return captured
