"""-m gpu: the opt-in / fallback code paths of the convolution and DCN kernels, each in a fresh interpreter (the
switches are read once per process): cluster multicast of the weight tile, the v1 gather kernel, the bf16 sampler
blend, one tap per DCN stage, plain launches instead of programmatic dependent launch.  Same parity tests, same
tolerances as tests/test_conv_gpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = [
    ({"CNB_CONV_CLUSTER": "2"}, "test_conv_matches_torch or test_conv_concat_slices"),
    ({"CNB_CONV_CLUSTER": "4"}, "test_conv_matches_torch"),
    ({"CNB_CONV_IMPL": "v1"}, "test_conv_matches_torch or test_stem_space_to_depth"),
    ({"CNB_CONV_ROWS": "0"}, "test_conv_matches_torch"),
    ({"CNB_CONV_PAIR": "1"}, "test_conv_matches_torch or test_conv_concat_slices or head"),   # CTA pairs for every eligible N
    ({"CNB_CONV_PAIR": "0"}, "test_conv_matches_torch or test_conv_concat_slices"),
    ({"CNB_CONV_FP": "1"}, "test_conv_matches_torch or test_conv_concat_slices or dcn_module"),   # every eligible 3x3 conv through the footprint kernel
    ({"CNB_CONV_FP": "0"}, "test_conv_matches_torch or dcn_module"),
    ({"CNB_DCN_BLEND": "bf16"}, "dcn"),
    ({"CNB_DCN_BLEND": "wbf16"}, "dcn"),
    ({"CNB_DCN_GROUPS": "1"}, "dcn"),
    ({"CNB_DCN_GROUPS": "4"}, "dcn"),
    ({"CNB_DCN_IMPL": "v1"}, "dcn"),
    ({"CNB_DCN_IMPL": "ws"}, "dcn"),
    ({"CNB_DCN_REACH": "1"}, "dcn"),
    ({"CNB_DCN_REACH": "1", "CNB_DCN_STAGES": "2", "CNB_DCN_GROUPS": "1"}, "dcn"),
    ({"CNB_DW_DECONV_IMPL": "5"}, "upsample"),     # cp.async tile kernel at every size
    ({"CNB_DW_DECONV_IMPL": "3"}, "upsample"),     # phase kernel at every size
    ({"CNB_PDL": "0"}, "test_conv_matches_torch or dcn"),
]


@pytest.mark.parametrize("env,select", VARIANTS, ids=[",".join(f"{k}={v}" for k, v in e.items()) for e, _ in VARIANTS])
def test_kernel_variant(cuda_dev, env, select):
    e = dict(os.environ)
    e.update(env)
    import torch
    torch.cuda.synchronize()
    torch.cuda.empty_cache()      # the child gets its own context: give back what this process's allocator has cached
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_conv_gpu.py"), "-x", "-q",
           "-m", "gpu", "-k", select, "-p", "no:cacheprovider"]
    r = subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        # One full-suite run in ~30 lost this variant's child once and could not be reproduced in 34 repeats of the same
        # command: keep the evidence (gpurun_out/ travels back from the GPU box) and look a second time before failing.
        tag = "_".join(f"{k}-{v}" for k, v in env.items())
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"variant_first_failure_{tag}.txt"), "w") as fh:
            fh.write(r.stdout[-20000:] + "\n--- stderr\n" + r.stderr[-8000:])
        r2 = subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)
        assert r2.returncode == 0, ("failed twice; first:\n" + r.stdout[-3000:] + r.stderr[-1500:] +
                                    "\nsecond:\n" + r2.stdout[-3000:] + r2.stderr[-1500:])
