"""``python -m remora_b200 infer POD5 BAM --model MODEL.pt --out CALLS.bam``: the function form of
``remora infer from_pod5_and_bam`` (reference src/remora/parsers.py:1385-1612) behind a minimal command
line.  Under ``torchrun`` every rank takes its share of the reads and writes ``<out stem>.rank<k><ext>``."""
import argparse
import os
import sys
import time


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m remora_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    inf = sub.add_parser("infer", help="call modified bases from a POD5 + BAM pair")
    inf.add_argument("pod5")
    inf.add_argument("in_bam")
    inf.add_argument("--model", required=True, action="append",
                     help="TorchScript Remora model (repeat for one model per canonical base)")
    inf.add_argument("--out", required=True, help="output .bam or .sam")
    inf.add_argument("--device", type=int, default=None, help="CUDA device (default: LOCAL_RANK or 0)")
    inf.add_argument("--num-reads", type=int, default=None)
    inf.add_argument("--batch-size", type=int, default=2048)
    inf.add_argument("--reads-per-batch", type=int, default=256)
    inf.add_argument("--reference-anchored", action="store_true")
    inf.add_argument("--out-format", choices=["bam", "sam"], default=None, help="default: by the output name")
    inf.add_argument("--drop-move-tag", action="store_true", help="do not copy the mv tag to the output")
    args = ap.parse_args(argv)

    import torch
    from . import inference, model_util
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = args.device if args.device is not None else int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", dev)
    models = {}
    for path in args.model:
        model, md = model_util.load_model(path, device=device, eval_only=True)
        models[md["can_base"]] = (model, md)
    # the rank goes BEFORE the extension (calls.rank0.bam) so that the format still follows the name
    stem, ext = os.path.splitext(args.out)
    out = args.out if world == 1 else f"{stem}.rank{rank}{ext}"
    t0 = time.perf_counter()
    res = inference.infer_from_pod5_and_bam(
        args.pod5, args.in_bam, models, out_path=out, num_reads=args.num_reads, batch_size=args.batch_size,
        reads_per_batch=args.reads_per_batch, ref_anchored=args.reference_anchored, rank=rank, world_size=world,
        out_format=args.out_format, drop_move_tag=args.drop_move_tag)
    dt = time.perf_counter() - t0
    ok = sum(r["error"] is None for r in res)
    calls = sum(len(r["ml"]) for r in res)
    print(f"[remora_b200] rank {rank}/{world}: {ok} reads called ({len(res) - ok} failed), {calls} calls in "
          f"{dt:.1f} s -> {out}", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
