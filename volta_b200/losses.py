"""Visual pre-training targets (reference: volta/losses.py:127-147).

`pre_vis_targets` maps the target id used in `config.visual_target_weights` to the width of its
prediction head.  The criteria themselves run on device (csrc/loss.cu):
  "0" kl_1601   - masked-region KL against the detector's 1601-way distribution (volta/losses.py:16-22)
  "3" xent_1600, "4" xent_400, "6" xent_1601 - confidence-weighted cross-entropies (volta/losses.py:87-133)
  "1" mse_2048, "5" huber_2048 - masked-region feature regression (volta/losses.py:25-33, 110-122)
Target "2" nce_2048 (:36-84) draws its negatives with the host RNG inside the criterion; no shipped config enables it and
the engine raises NotImplementedError for it.
"""
pre_vis_targets = {"0": 1601, "1": 2048, "2": 2048, "3": 1600, "4": 400, "5": 2048, "6": 1601}
pre_vis_criterion_names = {"0": "kl_1601", "1": "mse_2048", "2": "nce_2048", "3": "xent_1600", "4": "xent_400",
                           "5": "huber_2048", "6": "xent_1601"}
DEVICE_CRITERIA = {"0", "1", "3", "4", "5", "6"}
