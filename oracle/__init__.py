"""CPU oracle for the SR-inference hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic the reference delegates to
TensorFlow/Keras 2.10 (conv nets, ``tf.nn.depth_to_space``, ``tf.image.psnr/ssim``) and
to OpenCV (``cv2.resize(..., INTER_CUBIC)``), each function citing the reference
file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package never does, and it has no CPU fallback.

Pinning status:
* bicubic   - PINNED: checked against the reference's actual dependency call,
              ``cv2.resize`` (OpenCV 4.13.0 in this image), float32 and uint8.
* Lanczos-4 / bilinear / area - PINNED: ``cv2.resize`` outputs in tests/golden/resize_cv2.npz (the bilinear and area
              modes are checked on the device only; Lanczos-4 has a numpy restatement in bicubic.py).
* tiling    - PINNED: checked against ``SRModels/loading_methods.py::add_padding`` imported
              from /root/reference when generating ``tests/golden``, and against the dataset
              shape print-outs of the reference notebooks.
* conv nets - parity unpinned by execution (TensorFlow is not installable here); anchored on
              the architecture known-answers of the reference notebooks (parameter counts) and
              cross-checked fp32 vs fp64 and torch-conv vs explicit numpy im2col.
* PSNR/SSIM - parity unpinned by execution (tf.image absent); anchored on analytic
              known-answers listed in SURVEY.md section 8c and cross-checked against an
              independent scipy.ndimage evaluation of the same published definition
              (tests/test_oracle_metrics.py).
* skimage-style PSNR/SSIM (the classical benchmark notebook's metric definitions) - parity unpinned (scikit-image is
              absent): restated on the scipy.ndimage.uniform_filter primitive skimage itself calls, and checked
              against a direct evaluation of every valid window with numpy's sample (co)variance.
"""
