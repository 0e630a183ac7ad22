"""CPU: the plain-torch restatements in oracle/torch_oracle.py against golden values produced by the REAL reference
functions (tests/golden/make_golden_losses.py -> utils/slam_utils.py:91-165).  The GPU tests then compare the fused
kernels with the same golden file and with these restatements at full size."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import torch_oracle as TO  # noqa: E402
from test_losses import _run_golden_case  # noqa: E402

pytestmark = []


@pytest.mark.parametrize("name", ["mapping", "mapping_init", "tracking"])
def test_torch_restatement_matches_real_reference_golden(name):
    fn_map = lambda image, depth, gt_image, gt_depth, **kw: TO.reference_mapping_loss(image, depth, gt_image, gt_depth, **kw)
    _run_golden_case(name, torch.device("cpu"), fn_map, TO.reference_tracking_loss)
