"""CPU oracle for the EOT patch-attack hot path -- TEST INFRASTRUCTURE ONLY.

This package is an op-for-op restatement, in float32 NumPy, of what the reference
computes on its patch-attack hot path:

  * ``Patcher`` / ``Masker``            /root/reference/attacker.py:344-498,
                                        /root/reference/attack_detection.py:321-498
  * ``BrightnessMatcher``               /root/reference/brightness_matcher.py:25-73
  * the person-score objective          /root/reference/attacker.py:69-89,118-141,190-193,
                                        automl/efficientdet/tf2/postprocess.py:67-156,
                                        automl/efficientdet/tf2/anchors.py:30-58,117-165
  * ``tape.gradient(loss,[scale,patch])``  attacker.py:217 (chain of SURVEY.md section 3.2)

PARITY UNPINNED for Patcher / BrightnessMatcher / Masker / objective: the reference
holds no test, golden vector or fixture for them and the
arithmetic lives in third-party wheels that are absent here and on the GPU box
(tensorflow==2.8.1, tensorflow-addons==0.17.0, requirements.txt:4,16).  The TF/TFA op
semantics (ScaleAndTranslate, ImageProjectiveTransformV3 and its registered gradient,
rgb_to_yuv/yuv_to_rgb tables) are restated in ``oracle/tfops.py`` from the published
algorithms of those pinned versions.  What *is* pinned: the anchor arithmetic
(``tf2/postprocess_test.py:27-35,229`` first anchor) and the centring/clamping logic
of ``create`` against the importable ``adv_patch.AdversarialPatch._create``
(fixtures under ``tests/golden/``, generator ``tests/golden/make_golden.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product
(``mladversarialobjectdetection_b200``) never does, and fails loudly when its CUDA
library is missing.
"""
