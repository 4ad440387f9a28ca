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
  * first-pass NMS (``nms.py``), input pipeline (``input_pipeline.py``), uint8 inference twin (``adv_patch_u8.py``)

PARITY STATUS.  The arithmetic of the leaf TF/TFA kernels lives in third-party wheels that are absent here and on
the GPU box (tensorflow==2.8.1, tensorflow-addons==0.17.0, requirements.txt:4,16): ScaleAndTranslate,
ImageProjectiveTransformV3 and its registered gradient, rgb_to_yuv/yuv_to_rgb, NonMaxSuppressionV5 are restated
in ``oracle/tfops.py`` / ``oracle/nms.py`` from the published algorithms of those versions and stay UNPINNED, as
does the backward chain (no reference test, golden vector or fixture exists for it).

What IS pinned, bit for bit, by fixtures generated in the build container from the reference's own code
(``tests/golden/make_golden.py``):
  * ``Patcher.call``, ``BrightnessMatcher.call``, ``Masker.call`` (both branches): the reference's Python --
    control flow, expression order, casts, pad / where / clip / scatter sequence, box filter -- executed verbatim
    on a NumPy stand-in of the TF ops (``tests/golden/tf_numpy_shim.py``; its leaf kernels are the restated ones)
  * ``adv_patch.AdversarialPatch`` (``_create``, ``print_patch``, ``add_adv_to_img``): real NumPy + OpenCV run
  * ``train_data_generator.DataSequence._map_fn``: real NumPy + OpenCV run
  * ``tf2/anchors.py:Anchors``: real NumPy arithmetic; ``automl/efficientdet/nms_np.py`` for the NMS selections
  * the anchor known answer of ``tf2/postprocess_test.py:27-35,229``

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product
(``mladversarialobjectdetection_b200``) never does, and fails loudly when its CUDA
library is missing.
"""
