"""Runs the reference-side ctypes stub printed in INTEGRATION.md section 2 as it stands there (extracted from the markdown)
against the built library on a GPU: register a chunked light curve, batched lnlike, 50 chain steps through gprot_chain_run,
last-step log-posterior against the oracle.  usage (GPU box): python tools/check_integration_stub.py"""
import os
import re
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
code = re.search(r"```python\n(# gprot/b200.py.*?)```", md, re.S).group(1)
code = code.replace('ctypes.CDLL("libgprot_b200.so")', 'ctypes.CDLL(%r)' % os.path.join(ROOT, "gprotation_b200", "libgprot_b200.so"))
ns = {}
exec(compile(code, "INTEGRATION.md:gprot/b200.py", "exec"), ns)

import numpy as np  # noqa: E402

from gprotation_b200.lc import LightCurve  # noqa: E402
from gprotation_b200.model import GPRotModel  # noqa: E402
from gprotation_b200.synth import synthetic_lightcurve  # noqa: E402
from oracle import gp_oracle as o  # noqa: E402

t, y, yerr, truth = synthetic_lightcurve(3, 600, kind="harmonic")
lc = LightCurve(t, y, yerr, name="s", chunksize=300)
eng = ns["Engine"](0)
lid = eng.register(lc)
mod = GPRotModel(lc)                  # only its prior settings are used here
start = mod.sample_from_prior(32, seed=4)
lp = np.array([o.lnprior(x) for x in start])
ll = eng.lnlike(start, lid)
chunks = o.split_chunks(600, 300)
assert abs(ll[0] - o.lnlike(start[0], t, y, yerr, chunks)) <= 1e-9 * abs(ll[0])
chain, lnp, acc = eng.run_chain(mod, lid, start, lp, ll, 50, seed=7)
ref = o.lnlike(chain[-1, 5], t, y, yerr, chunks) + o.lnprior(chain[-1, 5])
assert abs(lnp[-1, 5] - ref) <= 1e-9 * abs(ref)
print("integration stub ok: chain", chain.shape, "acceptance %.2f" % acc.mean(), "last-step lnpost rel err %.1e" % abs((lnp[-1, 5] - ref) / ref))
