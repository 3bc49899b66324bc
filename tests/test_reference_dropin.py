"""The UNMODIFIED reference next to the CUDA path on the GPU box: cross-decode through
the stock decoder, and the drop-in swap of INTEGRATION.md under the reference's own integration tests.  Needs
the verbatim copy that tools/install_reference.py puts under baseline/_ref/ (or /root/reference)."""
import json
import os
import subprocess
import sys

import pytest

from oracle.load_reference import reference_available

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not reference_available(), reason="no reference tree (run tools/install_reference.py)")]

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def report():
    out = subprocess.run([sys.executable, "-W", "ignore", os.path.join(HERE, "dropin_driver.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_stock_decoder_reads_gpu_streams_and_vice_versa(report):
    for c in report["cross_decode"]:
        # the decoders agree to 1 LSB on either side's streams (fp32 inverse transform vs float64)
        assert c["gpu_stream_stock_vs_gpu_decoder_maxdiff"] <= 1, c
        assert c["stock_stream_stock_vs_gpu_decoder_maxdiff"] <= 1, c
    # smooth content without exact ties: the streams themselves are the reference's
    assert sum(c["streams_equal"] for c in report["cross_decode"]) >= len(report["cross_decode"]) - 1, report["cross_decode"]
    assert report["container"]["stock_vs_gpu_decode_maxdiff"] <= 1, report["container"]


def test_reference_integration_tests_pass_over_the_c_abi_stub(report):
    it = report["integration_tests"]
    assert (it["ran"], it["failures"], it["errors"]) == (5, 0, 0), it
    assert report["swapped_facade"]["decode_vs_stock_maxdiff"] <= 1, report["swapped_facade"]
    assert report["bad_rle"] == "(0, 19, 146880)", report["bad_rle"]
