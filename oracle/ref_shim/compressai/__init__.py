"""TEST INFRASTRUCTURE ONLY -- not product code.

Minimal pure-torch restatement of the CompressAI 1.2.6 symbols that
/root/reference/MLIC++ imports (requirements.txt:31 pins compressai==1.2.6; the
package is not vendored and cannot be installed here).  It exists so that the
*unmodified* reference model code can be imported in the build container to
generate golden vectors (oracle/make_golden.py).  It never travels into the
product path and is never imported on the GPU box.

Restated from the package's published behaviour; call sites in the reference:
  models/mlicpp.py:5-7,13-15,36,96-98,205-216,461-475
  modules/layers/res_blk.py:4,76,107,110-111 ; modules/layers/conv.py:5,32
  modules/transform/synthesis.py:4,21,25,67 ; utils/ckbd.py:3,75,82-90,128-129
"""
__version__ = "1.2.6-shim"
