"""CPU tests of host-side logic that prepares operands for the CUDA kernels (no compute calls into the library)."""
import pytest
import torch
import torch.nn.functional as F

from centernet_pytorch_lightning_b200 import ops


@pytest.mark.parametrize("K", [3, 5, 7])
def test_stem_space_to_depth_filter(K):
    """ops.stem_s2d_filter: the K x KX filter over pixel pairs reproduces the K x K convolution exactly
    (same products, same sums) -- checked with plain torch on the re-viewed input."""
    g = torch.Generator().manual_seed(K)
    B, H, W, Co = 2, 10, 12, 16
    w = torch.randn(Co, 3, K, K, generator=g, dtype=torch.float64)
    x = torch.randn(B, 3, H, W, generator=g, dtype=torch.float64)
    ws, (K_, KX, D) = ops.stem_s2d_filter(w.float())
    assert K_ == K and KX == 2 * D + 1 and ws.shape == (2 * Co, 8, K, KX)
    assert torch.count_nonzero(ws) == torch.count_nonzero(w.float()) * 2      # every tap once per output parity
    x4 = torch.zeros(B, H, W, 4, dtype=torch.float64)
    x4[..., :3] = x.permute(0, 2, 3, 1)
    xs = x4.view(B, H, W // 2, 8).permute(0, 3, 1, 2)
    y = F.conv2d(xs, ws.double(), padding=(K // 2, D)).permute(0, 2, 3, 1).reshape(B, H, W, Co)
    ref = F.conv2d(x, w.float().double(), padding=K // 2).permute(0, 2, 3, 1)
    assert torch.allclose(y, ref, atol=1e-12)


def test_standalone_probes_compile_for_sm100a(tmp_path):
    """tools/*.cu (tcgen05.mma issue-rate probe, TMA feed probe) are standalone nvcc programs against csrc/umma.cuh: they
    must keep compiling for sm_100a as the PTX wrappers change (nvcc cross-compiles without a GPU)."""
    import glob
    import os
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    srcs = sorted(glob.glob(os.path.join(root, "tools", "*.cu")))
    assert any(s.endswith("mma_probe.cu") for s in srcs)
    for src in srcs:
        out = tmp_path / (os.path.basename(src) + ".o")
        r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-c", src, "-o", str(out),
                            "-I", os.path.join(root, "centernet_pytorch_lightning_b200", "csrc"), "-I", os.path.join(root, "include")],
                           capture_output=True, text=True)
        assert r.returncode == 0, f"{src}:\n{r.stderr[-2000:]}"
