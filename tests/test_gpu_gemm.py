"""GPU: osd_gemm against torch over operand layouts, outputs, ragged M, split-K and the fused QKV epilogue
(tools/gemm_check.py), with the CTA-pair (cta_group::2) kernel at its default rule, switched off and forced on.  The switch
is read once per process, hence the subprocesses.  Replaces the reference's 1x1 Conv1d / Linear call sites
(common/attn.py:68-69 qkv_proj / out_proj, common/swiglu.py:21-24 proj_vg / proj_o, models/diffusion/backbone.py:63 proj_cl;
SURVEY.md section 2.1)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('pair', [None, '0', '1'])
def test_gemm_matches_torch(pair):
    env = dict(os.environ)
    env.pop('OSD_GEMM_PAIR', None)
    if pair is not None:
        env['OSD_GEMM_PAIR'] = pair
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'gemm_check.py')], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
    assert 'failures 0' in r.stdout
