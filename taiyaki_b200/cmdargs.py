"""Argument types of the entry points that the reference keeps in taiyaki/cmdargs.py: here the
boolean flag pairs (`AutoBool`, cmdargs.py:90-127), which the scripts under bin/ and misc/ share so
that the reference's command lines parse unchanged."""
import argparse
import os
import re


class AutoBool(argparse.Action):
    """`--flag` sets True, `--no-flag` sets False; neither takes a value, so a positional
    argument may follow either.  A default is required and named in the help text."""

    def __init__(self, option_strings, dest, default=None, required=False, help=None):
        if default is None:
            raise ValueError('You must provide a default with AutoBool action')
        if len(option_strings) != 1 or not option_strings[0].startswith('--'):
            raise ValueError('AutoBool takes a single option string prefixed with --')
        name = option_strings[0][2:]
        pair = ['--' + name, '--no-' + name]
        super().__init__(pair, dest, nargs=0, const=None, default=default, required=required,
                         help='{} (Default: {})'.format(help or '', pair[0] if default else pair[1]))

    def __call__(self, parser, namespace, values, option_string=None):
        setattr(namespace, self.dest, not option_string.startswith('--no-'))


class _Checked:
    """An argparse `type`: converts with `mytype`, then requires `accepts(value)`."""
    description = 'value'

    def __init__(self, mytype):
        self.mytype = mytype

    def accepts(self, value):
        return True

    def __repr__(self):
        return '{} {}'.format(self.description, self.mytype)

    def __call__(self, text):
        value = self.mytype(text)
        if not self.accepts(value):
            raise argparse.ArgumentTypeError('Argument must be {}'.format(self))
        return value


class Positive(_Checked):
    """Strictly positive values of `mytype` (cmdargs.py:207-225)."""
    description = 'positive'

    def accepts(self, value):
        return value > 0


class Bounded(_Checked):
    """Values of `mytype` in [lower, upper]; either bound may be left out (cmdargs.py:154-195)."""

    def __init__(self, mytype, lower=None, upper=None):
        super().__init__(mytype)
        assert lower is not None or upper is not None
        assert lower is None or upper is None or lower <= upper
        self.lower, self.upper = lower, upper

    def __repr__(self):
        return '{} in range [{}, {}]'.format(self.mytype, '-inf' if self.lower is None else self.lower,
                                             'inf' if self.upper is None else self.upper)

    def accepts(self, value):
        return (self.lower is None or value >= self.lower) and (self.upper is None or value <= self.upper)


def NonNegative(mytype):
    """Values of `mytype` that are >= 0 (cmdargs.py:198-204)."""
    return Bounded(mytype, lower=mytype(0))


def proportion(text):
    """A float in [0, 1] (cmdargs.py:228-230)."""
    return Bounded(float, 0.0, 1.0)(text)


class Maybe:
    """The string 'None', or whatever `mytype` accepts (cmdargs.py:130-151)."""

    def __init__(self, mytype):
        self.mytype = mytype

    def __repr__(self):
        return 'None or {}'.format(self.mytype)

    def __call__(self, text):
        if text == 'None':
            return None
        try:
            return self.mytype(text)
        except Exception:
            raise argparse.ArgumentTypeError('Argument must be {}'.format(self))


class FileExists(argparse.Action):
    """The named path must exist (cmdargs.py:28-36)."""

    def __call__(self, parser, namespace, values, option_string=None):
        if not os.path.exists(values):
            raise RuntimeError("File/path for '{}' does not exist, {}".format(self.dest, values))
        setattr(namespace, self.dest, values)


class FileAbsent(argparse.Action):
    """The named path must not exist yet (cmdargs.py:39-46)."""

    def __call__(self, parser, namespace, values, option_string=None):
        if os.path.exists(values):
            raise RuntimeError("File/path for '{}' exists, {}".format(self.dest, values))
        setattr(namespace, self.dest, values)


class DeviceAction(argparse.Action):
    """A device given as the scheduler or the user writes it (cmdargs.py:277-306): '2' and 'cuda2'
    become the integer 2 (what torch.device takes for GPU 2); anything else ('cuda:2', 'cuda',
    'cpu') is left for torch.device to interpret."""

    def __call__(self, parser, namespace, value, option_string=None):
        setattr(namespace, self.dest, self.convert(value))

    @staticmethod
    def convert(value):
        if value is None:
            return 'cpu'
        for pattern in (r'[0-9]+', r'cuda([0-9]+)'):
            m = re.match(pattern, value)
            if m:
                return int(m.group(m.lastindex or 0))
        return value
