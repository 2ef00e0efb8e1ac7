"""CPU suite: the import block of the reference's training / validation scripts resolves against shims/ with zero
edits (nzdownscale/downscaler/train.py:12,19-26,262,370,648; validate_ERA.py:6-7), and what it resolves to is this
package."""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REFERENCE_IMPORTS = '''
import lab as B
import torch
import deepsensor.torch
from deepsensor.data.loader import TaskLoader
from deepsensor.model.convnp import ConvNP
from deepsensor.train.train import train_epoch, set_gpu_default_device
from deepsensor.data.task import Task
from deepsensor.data import construct_circ_time_ds          # validate_ERA.py:7 (imported, not called on the hot path)
from deepsensor.data.processor import DataProcessor         # preprocess.py:23
from deepsensor.data.utils import construct_x1x2_ds         # preprocess.py:24
import deepsensor, deepsensornz_b200 as P
assert ConvNP is P.ConvNP and train_epoch is P.train_epoch and Task is P.Task
assert set_gpu_default_device is P.set_gpu_default_device
from deepsensornz_b200.loader import TaskLoader as TL
assert TaskLoader is TL
m = ConvNP(dim_yc=(3, 6, 1, 1), dim_yt=1, dim_aux_t=5, internal_density=50, encoder_scales=(0.01,) * 4,
           decoder_scale=0.02, verbose=False)
n = deepsensor.backend.nps.num_params(m.model)               # train.py:262
assert n == sum(p.numel() for p in m.model.parameters()) and n > 1_000_000
x = B.to_numpy(torch.tensor(1.5, dtype=torch.float64))       # train.py:370
assert float(x) == 1.5
s = B.sigmoid(torch.zeros(3))                                # train.py:648
assert torch.allclose(s, torch.full((3,), 0.5))
try:
    construct_circ_time_ds(None, None)
    raise SystemExit("ETL helper should refuse")
except ImportError as e:
    assert "hot path" in str(e)
print("SHIMS_OK", n)
'''


def test_reference_import_block_runs_against_shims():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "shims"), ROOT, env.get("PYTHONPATH", "")])
    out = subprocess.run([sys.executable, "-c", REFERENCE_IMPORTS], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SHIMS_OK 1145031" in out.stdout      # Cin = 15: 1 145 026 UNet + MLP parameters and the 5 frozen length scales


def test_lab_shim_numpy_inputs():
    sys.path.insert(0, os.path.join(ROOT, "shims"))
    try:
        import lab as B
        assert np.allclose(B.sigmoid(np.array([0.0, 100.0])), [0.5, 1.0])
        assert isinstance(B.to_numpy([torch.ones(2)])[0], np.ndarray)
    finally:
        sys.path.pop(0)
        sys.modules.pop("lab", None)
