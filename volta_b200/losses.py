"""Visual pre-training targets (reference: volta/losses.py:127-147).

`pre_vis_targets` maps the target id used in `config.visual_target_weights` to the width of its
prediction head.  The criteria themselves run on device (csrc/loss.cu):
  "0" kl_1601   - masked-region KL against the detector's 1601-way distribution (volta/losses.py:16-22)
  "3" xent_1600, "4" xent_400, "6" xent_1601 - confidence-weighted cross-entropies (volta/losses.py:87-133)
Targets "1" mse_2048, "2" nce_2048 and "5" huber_2048 are feature-regression objectives that no shipped
ctrl_* config enables; the engine raises NotImplementedError for them.
"""
pre_vis_targets = {"0": 1601, "1": 2048, "2": 2048, "3": 1600, "4": 400, "5": 2048, "6": 1601}
pre_vis_criterion_names = {"0": "kl_1601", "1": "mse_2048", "2": "nce_2048", "3": "xent_1600", "4": "xent_400",
                           "5": "huber_2048", "6": "xent_1601"}
DEVICE_CRITERIA = {"0"}
