"""Backbone plugin registry with the reference's contract (CenterNet/models/__init__.py:6-19):
`create_model("dla_34" | "resdcn_50" | "res_18" ...)` -> nn.Module with `.out_channels` and `forward(x) -> [Tensor]`.
`hourglass` is Hourglass-104 (large_hourglass.py:322-343; two stacks, inference schedule)."""
from .backbones.large_hourglass import get_large_hourglass_net
from .backbones.pose_dla_dcn import get_pose_net as get_dla_dcn
from .backbones.resnet import get_pose_net, get_pose_net_dcn

_model_factory = {
    "res": get_pose_net,         # default ResNet with dense deconvs (msra_resnet.py)
    "dla": get_dla_dcn,
    "resdcn": get_pose_net_dcn,  # ResNet + DCN + deconv (resnet_dcn.py)
    "hourglass": get_large_hourglass_net,
}


def create_model(arch):
    num_layers = int(arch[arch.find("_") + 1:]) if "_" in arch else 0
    arch = arch[: arch.find("_")] if "_" in arch else arch
    if arch not in _model_factory:
        raise KeyError(f"backbone '{arch}' is not built in centernet_b200; available: {sorted(_model_factory)}")
    return _model_factory[arch](num_layers=num_layers)
