#!/usr/bin/env python3
"""Dump the JSON representation of a model -- taiyaki's bin/dump_json.py (:11-38): the structured
description `model.json()` of a checkpoint (layer types, sizes, parameters in Guppy's layout) plus the
md5 sum of the checkpoint file.  Host only: loading a checkpoint and describing it needs no GPU.

    dump_json.py [--output model.json] model.checkpoint
"""
import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from taiyaki_b200 import helpers  # noqa: E402
from taiyaki_b200.cmdargs import FileAbsent, FileExists  # noqa: E402
from taiyaki_b200.json import JsonEncoder  # noqa: E402


def get_parser():
    p = argparse.ArgumentParser(description='Dump JSON representation of model',
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--output', default=None, metavar='filename', action=FileAbsent, help='Write output to file')
    p.add_argument('model', action=FileExists, help='Model checkpoint')
    return p


def file_md5(filename, nblock=1024):
    """taiyaki/helpers.py:302-317."""
    hasher = hashlib.md5()
    with open(filename, 'rb') as fh:
        for block in iter(lambda: fh.read(nblock * hasher.block_size), b''):
            hasher.update(block)
    return hasher.hexdigest()


def main(argv=None):
    args = get_parser().parse_args(argv)
    json_out = helpers.load_model(args.model).json()
    json_out['md5sum'] = file_md5(args.model)
    fh = sys.stdout if args.output is None else open(args.output, 'w')
    try:
        json.dump(json_out, fh, indent=4, cls=JsonEncoder)
    finally:
        if fh is not sys.stdout:
            fh.close()
    return json_out


if __name__ == '__main__':
    main()
