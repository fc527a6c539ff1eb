"""Load tests/golden/*.npz (written by scripts/make_golden.py from the live reference)."""
import ast
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["tiny_eval", "tiny_train", "color_mode4", "nobias_mu_param", "nobehav", "default_dims"]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.meta = ast.literal_eval(str(z["meta"]))
        self.args = self.meta["args"]
        self.sd = {k[3:]: z[k] for k in z.files if k.startswith("sd/")}
        self.mice = {}
        for m in self.meta["neurons"]:
            d = {k[len(m) + 1:]: z[k] for k in z.files if k.startswith(m + "/") and "/grad/" not in k}
            d["grads"] = {k[len(m) + 6:]: z[k] for k in z.files if k.startswith(m + "/grad/")}
            self.mice[m] = d

    def core_config(self):
        from oracle.v1t_oracle import CoreConfig

        c, h, w = self.meta["in_shape"]
        a = self.args
        return CoreConfig(in_ch=c, in_h=h, in_w=w, patch_size=a["patch_size"], patch_stride=a["patch_stride"],
                          emb_dim=a["emb_dim"], num_heads=a["num_heads"], mlp_dim=a["mlp_dim"],
                          num_blocks=a["num_blocks"], behavior_mode=a["behavior_mode"],
                          use_bias=not a["disable_bias"])


def rel_err(a, b):
    """max|a-b| / max|b| — the tolerance metric used throughout."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / denom
