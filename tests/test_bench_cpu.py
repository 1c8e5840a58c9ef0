"""bench.py's CPU legs (cpu_baseline, --impl reference) run the oracle alone: the product package and its shared library stay out
of that process, and the committed state_dict templates (oracle/shapes, dumped from the unmodified reference) give the same
seeded weights as the product's own parameter tree."""
import json
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def test_reference_arm_does_not_import_the_product():
    code = ("import sys, bench\n"
            "orc = bench.cpu_oracle('MLICPP_S')\n"
            "from oracle import weights\n"
            "mp = bench.cpu_step(orc, 'forward', weights.synthetic_image(1, 64, 64, seed=1))\n"
            "bad = [m for m in sys.modules if m.startswith('mlic_b200')]\n"
            "assert not bad, bad\n"
            "print(mp)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    assert float(r.stdout.strip()) == pytest.approx(64 * 64 / 1e6)


@pytest.mark.parametrize("name", ["MLICPP_L", "MLICPP_M_SMALL_DEC", "MLICPP_L_VBR", "MLICPP_S"])
def test_state_dict_template_matches_the_parameter_tree(name):
    import mlic_b200
    from oracle import weights
    a = weights.seeded_state_dict(mlic_b200.get_model(name).state_dict(), 1234, y_gain=4.0, sigma_spread=2.0)
    b = weights.reference_state_dict(name, 1234, y_gain=4.0, sigma_spread=2.0)
    assert not [k for k in b if k not in a]
    for k, v in b.items():
        assert a[k].shape == v.shape and torch.equal(a[k].to(v.dtype), v), k
    # what the template leaves out are only the derived tables update() fills
    assert all(k.rsplit(".", 1)[-1] in ("_quantized_cdf", "_offset", "_cdf_length", "scale_table") for k in a if k not in b)


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--config", "forward"],
                       capture_output=True, text=True, cwd=ROOT, env={**__import__("os").environ, "MLIC_BENCH_REF_SMALL": "1"})
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "MP/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
