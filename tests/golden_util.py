"""Shared helpers for tests that use the committed golden vectors (tests/golden/*.pt)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import make_golden  # noqa: E402  (case table + seeded batch; does not import the reference)
from volta_b200.config import BertConfig  # noqa: E402
from volta_b200.synthetic import seeded_state_dict  # noqa: E402

MINI_CASES = [n for n in make_golden.CASES if n.startswith("mini")]
FULL_CASES = [n for n in make_golden.CASES if n.startswith("full")]


def load_case(name):
    g = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), weights_only=False)
    cfg = BertConfig.from_dict(g["config"])
    batch = make_golden.case_batch(name)
    return g, cfg, batch


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))
