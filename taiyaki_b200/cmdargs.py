"""Argument types of the entry points that the reference keeps in taiyaki/cmdargs.py: here the
boolean flag pairs (`AutoBool`, cmdargs.py:90-127), which the scripts under bin/ and misc/ share so
that the reference's command lines parse unchanged."""
import argparse


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
