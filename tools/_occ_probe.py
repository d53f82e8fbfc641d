import os, sys
os.environ["FSB_DEBUG"] = "1"
sys.path.insert(0, ".")
import torch
from oracle import lm_oracle as O
from tests.lm_util import build_model
cfg = O.tiny_config()
m = build_model(cfg, O.make_weights(cfg, seed=1), debug=False)
print("ok")
