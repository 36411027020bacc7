"""patch_reference() against the real reference modules (build container only: /root/reference is
absent on the GPU box, where these tests skip).  Checks rebinding in the CONSUMER modules and that the
reference's own model classes then build mvs_b200 CostRegNets with identical state_dict keys."""
import os
import subprocess
import sys

import pytest

import cases

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")

SCRIPT = r'''
import sys, warnings, io, contextlib
warnings.filterwarnings("ignore")
sys.path.insert(0, {root!r}); sys.path.insert(0, {proj!r})
import torch
{pre}
import models
{imports}
ref_keys = {ref_keys}
import mvs_b200
done = mvs_b200.patch_reference()
{checks}
print("OK", sorted(done))
'''


def run(proj, pre, imports, ref_keys, checks):
    code = SCRIPT.format(root=cases.ROOT, proj=os.path.join(REF, proj), pre=pre, imports=imports, ref_keys=ref_keys,
                         checks=checks)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "OK" in p.stdout, p.stdout + p.stderr


def test_patch_mvsnet():
    run("MVSNet", "", "from models import mvsnet, module",
        "sorted(mvsnet.MVSNet(refine=False).state_dict())",
        "assert mvsnet.homo_warping is mvs_b200.ops.homo_warping and module.homo_warping is mvs_b200.ops.homo_warping\n"
        "assert mvsnet.MVSNet is mvs_b200.mvsnet.MVSNet and models.MVSNet is mvs_b200.mvsnet.MVSNet\n"
        "assert mvsnet.FeatureNet is mvs_b200.mvsnet.FeatureNet\n"
        "m = mvsnet.MVSNet(refine=False)\n"
        "assert type(m.cost_regularization).__module__ == 'mvs_b200.modules'\n"
        "assert sorted(m.state_dict()) == ref_keys")


def test_patch_cas():
    run("CasMVSNet", "", "from models import cas_mvsnet, module",
        "(lambda f: (lambda m: sorted(m.state_dict()))(cas_mvsnet.CascadeMVSNet()))(0)",
        "assert cas_mvsnet.DepthNet is mvs_b200.modules.DepthNet and cas_mvsnet.homo_warping is mvs_b200.ops.homo_warping\n"
        "buf = io.StringIO()\n"
        "with contextlib.redirect_stdout(buf): m = cas_mvsnet.CascadeMVSNet()\n"
        "assert type(m.cost_regularization[0]).__module__ == 'mvs_b200.modules' and type(m.DepthNet).__module__ == 'mvs_b200.modules'\n"
        "assert sorted(m.state_dict()) == ref_keys")


def test_patch_cvp():
    run("CVP-MVSNet", "import types", "from models import net, modules",
        "sorted(net.network(types.SimpleNamespace(nsrc=2, nscale=2, mode='test')).state_dict())",
        "assert net.proj_cost is mvs_b200.modules.proj_cost and net.homo_warping is mvs_b200.ops.homo_warping_cvp\n"
        "m = net.network(types.SimpleNamespace(nsrc=2, nscale=2, mode='test'))\n"
        "assert type(m.cost_reg_refine).__module__ == 'mvs_b200.modules'\n"
        "assert sorted(m.state_dict()) == ref_keys")
