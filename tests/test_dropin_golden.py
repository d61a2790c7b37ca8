"""BASELINE config 1 (CPU): decode sanity.bin with the reference's own parser -- with
OUR `sao` module swapped in for the reference's by name -- split the log with the
reference's tools/gen_logs.py and byte-compare all 95 files of test/golden.  Needs
/root/reference (skipped on the GPU box; the hashes of the golden files are committed in
tests/golden/golden_manifest.json so the skip is visible, not silent)."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

from conftest import GOLDEN, REPO
from oracle import refshim

RUNNER = r"""
import os, sys, types, runpy
sys.path.insert(0, {repo!r})
from oracle import refshim
import p265_b200
refshim.build()
refshim._stub_matplotlib()
sys.path.insert(0, refshim.SHIM_DIR)
if {swap!r}:
    sys.path.insert(0, p265_b200.dropin_path())       # our sao.py wins over the reference's
os.makedirs("logs", exist_ok=True)
import sao, dec
assert ("p265_b200" in sao.__file__) == bool({swap!r}), sao.__file__
args = types.SimpleNamespace(bitstream=os.path.join(refshim.SHIM_DIR, "sanity.bin"),
                             skip_syntax_dump=0, output=None, plot=None)
d = dec.Decoder(args)
try:
    d.decode()
except SystemExit:
    pass
import logging
logging.shutdown()
os.chdir("logs")
runpy.run_path(os.path.join(refshim.SHIM_DIR, "gen_logs.py"))
print("pictures", len(d.ctx.dpb.images))
"""


@pytest.mark.skipif(not refshim.reference_available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("swap", [True])
def test_golden_logs_byte_identical_with_our_sao_module(tmp_path, swap):
    script = tmp_path / "run.py"
    script.write_text(RUNNER.format(repo=REPO, swap=swap))
    r = subprocess.run([sys.executable, str(script)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert "pictures 3" in r.stdout, r.stdout[-500:] + r.stderr[-2000:]
    manifest = json.load(open(os.path.join(GOLDEN, "golden_manifest.json")))["files"]
    assert len(manifest) == 95
    bad = []
    for name, meta in manifest.items():
        p = tmp_path / "logs" / name
        data = p.read_bytes() if p.exists() else b""
        if hashlib.sha256(data).hexdigest() != meta["sha256"]:
            bad.append(name)
        ref = open(os.path.join(refshim.REF_ROOT, "test", "golden", name), "rb").read()
        assert hashlib.sha256(ref).hexdigest() == meta["sha256"], "manifest is stale: " + name
    assert not bad, "%d golden files differ: %s" % (len(bad), bad[:5])


def test_manifest_is_committed():
    m = json.load(open(os.path.join(GOLDEN, "golden_manifest.json")))
    assert m["regenerated_identical"] == 95 and not m["mismatch"]
