"""Backbone plugin registry with the reference's contract (CenterNet/models/__init__.py:6-19):
`create_model("dla_34")` -> nn.Module with `.out_channels` and `forward(x) -> [Tensor]`."""
from .backbones.pose_dla_dcn import get_pose_net as get_dla_dcn

_model_factory = {
    "dla": get_dla_dcn,
}


def create_model(arch):
    num_layers = int(arch[arch.find("_") + 1:]) if "_" in arch else 0
    arch = arch[: arch.find("_")] if "_" in arch else arch
    if arch not in _model_factory:
        raise KeyError(f"backbone '{arch}' is not built in centernet_b200 yet; available: {sorted(_model_factory)}")
    return _model_factory[arch](num_layers=num_layers)
