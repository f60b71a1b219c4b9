"""Writes config/*.json - the sub-layer schedules of the eight model configurations the reference ships
(config/*.json there) - from volta_b200.config.named_config_dict, so the files and the code cannot drift apart.
tests/test_config.py checks the round trip and, where /root/reference is mounted, equality with the reference's files."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from volta_b200.config import NAMED_CONFIGS, named_config_dict  # noqa: E402

os.makedirs(os.path.join(ROOT, "config"), exist_ok=True)
for name in NAMED_CONFIGS:
    with open(os.path.join(ROOT, "config", name + ".json"), "w") as f:
        json.dump(named_config_dict(name), f, indent=2)
    print("config/%s.json" % name)
