"""Model loading / saving and logging glue (the subset of taiyaki/helpers.py
the training entry point uses: load_model :82-136, save_model :32-79,
guess_model_stride :150-162, Logger :260-300, WindowedExpSmoother :212-257)."""
import copy
import importlib.util
import os
import sys

import numpy as np
import torch


def _load_python_model(model_file, **model_kwargs):
    """Load a `network(...)` factory from a model-definition file
    (helpers.py:17-29; importlib instead of the removed `imp`)."""
    spec = importlib.util.spec_from_file_location('netmodule', model_file)
    netmodule = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(netmodule)
    return netmodule.network(**model_kwargs)


class _legacy_rnn_pickles:
    """Whole-module pickles written by torch 1.x (the reference pins 1.5.1,
    requirements.txt:23; its shipped models/*.checkpoint are such files) hold
    nn.LSTM / nn.GRU objects without `_flat_weights`, which current torch's
    RNNBase.__setstate__ dereferences.  While a checkpoint is being loaded the
    missing list is rebuilt from the pickled parameter names."""

    def __enter__(self):
        base = torch.nn.RNNBase
        self.orig = orig = base.__setstate__

        def setstate(module, d):
            names = d.get('_all_weights')
            if '_flat_weights' not in d and names and isinstance(names[0][0], str):
                d = dict(d)
                flat = [n for layer in names for n in layer]
                d['_flat_weights_names'] = flat
                d['_flat_weights'] = [d['_parameters'].get(n) for n in flat]
            orig(module, d)
        base.__setstate__ = setstate

    def __exit__(self, *exc):
        torch.nn.RNNBase.__setstate__ = self.orig


def load_model(model_file, params_file=None, model_metadata=None, **model_kwargs):
    """Load a model from a .py definition or a .checkpoint pickle (helpers.py:82-136).
    Checkpoints pickled by the reference name their classes `taiyaki.layers.*`; they
    resolve through the `taiyaki` alias package of this repository."""
    if os.path.splitext(model_file)[1] == '.py':
        network = _load_python_model(model_file, **model_kwargs)
    else:
        import warnings
        import taiyaki  # noqa: F401  (registers taiyaki.layers -> taiyaki_b200.layers)
        with _legacy_rnn_pickles(), warnings.catch_warnings():
            warnings.simplefilter('ignore')       # SourceChangeWarning of every pickled class
            network = torch.load(model_file, map_location='cpu', weights_only=False)
    if params_file is not None:
        network.load_state_dict(torch.load(params_file, map_location='cpu'))
    if model_metadata is not None:
        network.metadata = model_metadata
    elif not hasattr(network, 'metadata'):
        network.metadata = {'reverse': False, 'standardize': True, 'version': 0}
    return network


def save_model(network, outdir, index=None, model_skeleton=None):
    """Write model_checkpoint_%05d.checkpoint / .params (helpers.py:32-79)."""
    basename = 'model_final' if index is None else 'model_checkpoint_{:05d}'.format(index)
    model_file = os.path.join(outdir, basename + '.checkpoint')
    params_file = os.path.join(outdir, basename + '.params')
    state = {k: v.detach().cpu() for k, v in network.state_dict().items()}
    torch.save(state, params_file)
    # pickle a CPU copy: moving the live network would detach the flat gradient views
    clone = copy.deepcopy(network)
    for p in clone.parameters():
        p.grad = None
    torch.save(clone.cpu(), model_file)
    return model_file, params_file


def get_model_device(model):
    return next(model.parameters()).device


def guess_model_stride(net):
    """Ratio of input to output time steps, measured on a dummy chunk
    (helpers.py:150-162).  The recurrent layers have no CPU path, so the probe
    runs on the model's device."""
    device = get_model_device(net)
    with torch.no_grad():
        out = net(torch.zeros((720, 1, 1), dtype=torch.float32, device=device))
    return int(round(720 / out.size()[0]))


class Logger(object):
    """Writes to a file and, unless quiet, to stdout (helpers.py:260-300)."""

    def __init__(self, log_file_name, quiet=False):
        try:
            self.fh = open(log_file_name, 'w', buffering=1)
        except Exception:
            sys.stderr.write("Failed to open log file {}\n".format(log_file_name))
            sys.exit(1)
        self.quiet = quiet

    def write(self, message):
        if not self.quiet:
            sys.stdout.write(message)
            sys.stdout.flush()
        try:
            self.fh.write(message)
        except IOError as e:
            print("Failed to write to log\n Message: {}\n Error: {}".format(message, repr(e)))


class WindowedExpSmoother(object):
    """Windowed exponential smoother for the reported loss (helpers.py:212-257)."""

    def __init__(self, alpha=0.95, n_vals=100):
        self.alpha = alpha
        self.weights = np.power(alpha, np.arange(n_vals))
        self.vals = np.full(n_vals, np.nan)
        self.n_valid_vals = 0
        self.n_vals = n_vals

    @property
    def value(self):
        if self.n_valid_vals == 0:
            return np.nan
        return np.average(self.vals[:self.n_valid_vals],
                          weights=self.weights[:self.n_valid_vals])

    def update(self, val):
        self.vals[1:] = self.vals[:-1]
        self.vals[0] = val
        self.n_valid_vals = min(self.n_valid_vals + 1, self.n_vals)
