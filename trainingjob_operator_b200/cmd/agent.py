"""Stand-alone node agent daemon (``python -m trainingjob_operator_b200.cmd.agent --master URL``)."""
from __future__ import annotations

import argparse
import os
import sys

from ..agent.agent import NodeAgent
from ..client.clientset import new_for_config
from ..signals import setup_signal_handler
from ..utils import klog
from .options import TrainingJobOperatorOption, resolve_master


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="aitj-agent")
    ap.add_argument("--master", default="")
    ap.add_argument("--kubeconfig", default="")
    ap.add_argument("--gpus", type=int, default=None)
    ap.add_argument("--workdir", default=os.path.expanduser("~/.aitj"))
    ap.add_argument("--image-map", action="append", default=[], help="image=command (repeatable)")
    ap.add_argument("--v", type=int, default=0)
    ap.add_argument("--warm-pool", type=int, default=-1,
                    help="parked pre-imported interpreters for fast replica start (-1 = one per GPU slot, 0 = off)")
    ap.add_argument("--gpu-visibility", default="pinned", choices=["pinned", "all"],
                    help="pinned: CUDA_VISIBLE_DEVICES holds only a replica's bound GPU; all: every GPU stays visible and "
                         "LOCAL_RANK names the bound one (workers that map their peers' memory need this)")
    args = ap.parse_args(argv)
    klog.configure(args.v, True)
    master = resolve_master(TrainingJobOperatorOption(master_url=args.master, kubeconfig=args.kubeconfig))
    import shlex

    image_map = {kv.split("=", 1)[0]: shlex.split(kv.split("=", 1)[1]) for kv in args.image_map if "=" in kv}
    from .local import _auto_pool

    agent = NodeAgent(new_for_config(master=master), num_gpus=args.gpus, workdir=args.workdir, image_map=image_map,
                      warm_pool=_auto_pool(args.warm_pool, args.gpus), gpu_visibility=args.gpu_visibility)
    stop = setup_signal_handler()
    agent.start(stop)
    print(f"aitj-agent up: {agent.num_gpus} GPU slot(s), master {master}", flush=True)
    stop.wait()
    agent.shutdown(kill=False)
    return 0


if __name__ == "__main__":
    sys.exit(main())
