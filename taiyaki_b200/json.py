"""JSON encoder for the structured layer descriptions `layers.*.json()` returns
(taiyaki/json.py:12-60): numpy scalars and arrays, torch parameters and tensors become plain
numbers and nested lists -- the Guppy-compatible model dump of bin/dump_json.py."""
import json

import numpy as np
import torch


class JsonEncoder(json.JSONEncoder):
    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        if isinstance(obj, torch.Tensor):           # parameters included
            return obj.detach().cpu().numpy().tolist()
        return super().default(obj)
