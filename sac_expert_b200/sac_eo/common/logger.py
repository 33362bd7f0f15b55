"""Dict logger with the reference's pickle schema {param, train, final}
(``/root/reference/sac_eo/common/logger.py:5-90``): append-merge into an existing checkpoint file."""
import os
import pickle

import numpy as np


class Logger:
    def __init__(self):
        self.reset()
        self.param_dict, self.final_dict = {}, {}

    def reset(self):
        self.train_dict = {}

    def log_train(self, kv):
        for k, v in kv.items():
            self.train_dict.setdefault(k, []).append(v)

    def log_params(self, kv):
        self.param_dict.update(kv)

    def log_final(self, kv):
        self.final_dict.update(kv)

    def dump(self):
        return {"param": self.param_dict, "train": {k: np.array(v) for k, v in self.train_dict.items()},
                "final": self.final_dict}

    def dump_and_save(self, log_path, log_name):
        out = self.dump()
        os.makedirs(log_path, exist_ok=True)
        fn = os.path.join(log_path, log_name)
        if os.path.isfile(fn):
            with open(fn, "rb") as f:
                old = pickle.load(f)
            for k, v in out["train"].items():
                if k in old["train"]:
                    out["train"][k] = np.concatenate([old["train"][k], v], 0)
            for k, v in old["train"].items():
                out["train"].setdefault(k, v)
        with open(fn, "wb") as f:
            pickle.dump(out, f)
        return out
