"""CPU checks of the drop-in import names: with compat.install() the UNMODIFIED reference
package imports and its CPU-constructible pieces build on our modules. Skipped where the reference tree
is not mounted (the GPU box)."""
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "vision3d")), reason="reference not mounted")


def test_reference_imports_resolve_to_the_dropins():
    import vision3d_b200.compat as compat
    compat.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import vision3d  # noqa: F401
    from vision3d.core import AnchorGenerator, Preprocessor, cfg
    from vision3d.detector import PV_RCNN, Second  # noqa: F401
    from vision3d.detector.proposal import ProposalLayer
    from vision3d.detector.roi_grid_pool import RoiGridPool
    from vision3d.detector.sparse_cnn import make_sparse_conv_layer, make_subm_layer
    from vision3d.ops import iou_nms
    assert iou_nms._C.__name__ == "vision3d_b200.compat.c_ext"
    assert iou_nms.box_iou_rotated.__module__ == "vision3d_b200.compat.c_ext"
    layer = make_subm_layer(4, 16, 3, indice_key="subm0")  # stray positional 3 lands in `stride`
    assert type(layer).__module__ == "vision3d_b200.compat.spconv" and layer[0].stride == [1, 1, 1]
    assert tuple(layer[0].weight.shape) == (3, 3, 3, 4, 16) and layer[0].bias is None
    conv = make_sparse_conv_layer(64, 64, (3, 1, 1), (2, 1, 1))
    assert conv[0].kernel_size == [3, 1, 1] and conv[0].stride == [2, 1, 1]
    pre = Preprocessor(cfg)  # builds spconv.utils.VoxelGenerator without needing a device
    assert pre.voxel_generator.grid_size.tolist() == [1408, 1600, 40]
    assert tuple(AnchorGenerator(cfg).anchors.shape) == (3, 2, 200, 176, 7)
    ProposalLayer(cfg)
    pool = RoiGridPool(cfg)
    assert type(pool.pnet).__module__ == "vision3d_b200.compat.pointnet2.pointnet2_modules"


def test_mirror_matches_reference_anchors_and_parameter_names():
    """vision3d_b200.second mirrors pieces of the reference that cannot travel to the GPU box; check them
    against the real thing here."""
    import torch
    import vision3d_b200.compat as compat
    compat.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from vision3d.core import AnchorGenerator, cfg
    from vision3d.core.box_encode import decode
    from vision3d_b200 import second
    mine = second.make_anchors(second.three_class_config())
    ref = AnchorGenerator(cfg).anchors
    assert torch.allclose(mine, ref, atol=1e-5), (mine - ref).abs().max()
    d, a = torch.randn(50, 7) * 0.1, ref.view(-1, 7)[:50]
    assert torch.allclose(second.decode_boxes(d, a), decode(d, a), atol=1e-6)
    # parameter names of the mirror == names the reference's Second would have (sparse_cnn.py / second.py)
    names = set(second.SecondB200(second.three_class_config()).state_dict().keys())
    for key in ["cnn.blocks.0.0.0.weight", "cnn.blocks.3.3.0.weight", "cnn.blocks.2.1.1.running_mean",
                "rpn.down_block.1.weight", "rpn.up_block.0.weight", "head.conv_cls.bias", "head.conv_reg.weight"]:
        assert key in names, key


def test_rpn_mirror_equals_reference_rpn_forward():
    """RPNB200 (the module the engine folds into cudnn conv+bias+ReLU calls) vs the reference's own RPN
    (detector/second.py:47-93) with the same state_dict, eval mode, on the CPU: identical outputs."""
    import torch
    import vision3d_b200.compat as compat
    compat.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from vision3d.detector.second import RPN
    from vision3d_b200 import second
    torch.manual_seed(11)
    ref = RPN().eval()
    with torch.no_grad():
        for m in ref.modules():  # non-trivial BN statistics
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.1)
    mine = second.RPNB200().eval()
    mine.load_state_dict(ref.state_dict())          # same parameter names and shapes
    x = torch.randn(1, 128, 20, 24)
    with torch.no_grad():
        assert torch.equal(mine(x), ref(x))
