"""cfg1 (BASELINE.json configs[0]): the reference's UNCHANGED NPP_completion/train.py, run for 20 iterations on the
bundled 211 x 325 example, once on the reference's own ``models`` package and once with this repo's drop-in ``models``
package shadowing it -- "scripts run unchanged" is the boundary's whole claim.  Both runs start
from the same seeds (np.random.seed(0), torch.random.manual_seed(0), train.py:16-17), draw the same pixels and patches
and must report the same losses up to the fp16-operand tolerance of the fused kernels.

The reference checkout travels to the GPU box as baseline/_ref (git-ignored copy made by __graft_entry__.build());
without it the test is skipped."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "run_reference_script.py")


def _run(impl, loss_type, iters=20, extra=()):
    env = dict(os.environ)
    env.pop("NPP_B200_EMBED", None)
    out = subprocess.run([sys.executable, RUNNER, "--impl", impl, "--iters", str(iters), "--loss_type", loss_type, *extra],
                         capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    if "unavailable" in rec:
        pytest.skip(rec["unavailable"])
    return rec


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("loss_type", ["l2", "robust_loss_adaptive"])
def test_unchanged_completion_script_on_dropin_models(loss_type):
    ref = _run("reference", loss_type)
    ours = _run("dropin", loss_type)
    assert "baseline" in ref["models_package"] or "reference" in ref["models_package"]
    assert ours["models_package"].endswith("_b200/models"), ours["models_package"]
    a, b = np.array(ref["losses"]), np.array(ours["losses"])
    assert len(a) == len(b) == 20, (len(a), len(b))
    assert np.isfinite(b).all()
    # same batches, same initial weights: the first losses agree to the kernels' tolerance; later iterations drift with
    # Adam's sign-like first steps (fp16-operand noise flips tiny gradients), so the whole trajectory gets a looser bound
    assert abs(a[0] - b[0]) <= 2e-3 * abs(a[0]), (a[0], b[0])
    assert np.abs(a - b).max() <= 0.05 * np.abs(a).max(), (a.tolist(), b.tolist())
    print(f"cfg1 {loss_type}: reference {ref['seconds']:.1f} s, drop-in {ours['seconds']:.1f} s for 20 iterations; "
          f"first loss {a[0]:.6f} / {b[0]:.6f}, last {a[-1]:.6f} / {b[-1]:.6f}")
