"""``trainingjob-operator`` binary entry point.

Parity: /root/reference/cmd/main.go:11-23 -- build options, bind flags, parse (klog flags included),
``app.Run``, fatal on error.  Usage mirrors README.md:11::

    python -m trainingjob_operator_b200.cmd.main --master 127.0.0.1:8001 --v 4 --thread-num 1000 \
        --logtostderr --leader-elect=true --enable-creating-failed=true
"""
from __future__ import annotations

import argparse
import sys

from ..utils import klog
from . import options, server


def main(argv=None) -> int:
    parser = argparse.ArgumentParser(prog="trainingjob-operator", description=__doc__,
                                     formatter_class=argparse.RawDescriptionHelpFormatter)
    options.add_flags(parser, options.new_training_job_operator_option())
    opt = options.from_args(parser.parse_args(argv))
    klog.configure(opt.v, opt.logtostderr)
    try:
        server.run(opt)
    except server.LeaseLost:
        return 255
    except Exception as e:  # noqa: BLE001
        klog.error("%r", e)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
