"""if_glow training throughput on synthetic data.

    python train_glow.py [--model if_glow_mnist|if_glow_cifar|if_glow_imagenet32] [--steps K] [--warmup W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P train_glow.py --model if_glow_imagenet32

One process per GPU; the batch is sharded over the ranks (weak scaling: the per-GPU batch is the
model's batch size) and the only communication is DistributedDataParallel's NCCL all-reduce of the
gradient buckets -- replacing the reference's single-process nn.DataParallel
(inf/if_multiGPU_imagenet32.py:410-411).  A step = forward (inverse-conv layers through the
sm_100a kernels), loss with NaN samples zeroed (inf/train/experiment.py:191-192), backward, gradient clipping,
Adam update under the linear learning-rate warm-up (experiment.py:197-202), as in the reference loop
(experiment.py:272-311) minus its per-layer print and parameter clamping.  Prints one JSON line (rank 0):
images/s of the whole job, max over ranks.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="if_glow_mnist")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--coupling-width", type=int, default=128)
    ap.add_argument("--groups", type=int, default=1)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--lr-warmup-steps", type=int, default=200,
                    help="linear warm-up of the learning rate over this many optimizer steps (experiment.py:197-202)")
    ap.add_argument("--init", default="reference", choices=["reference", "damped"],
                    help="inverse-conv weights: the reference initialisation (dirac at the kernel's spatial centre + noise: "
                         "for k = 3 that is the (1,1) tap, an operator whose inverse amplifies along the diagonal) or the "
                         "same with the non-centre taps scaled by 0.05 (a well-conditioned start)")
    ap.add_argument("--fused", action="store_true",
                    help="ActNorm (and each block's Squeeze) inside the inverse-conv kernels (layers.ActNormInvFlow)")
    ap.add_argument("--ddp", action="store_true", help="N > 1: torch DistributedDataParallel (eager) instead of the flat-bucket all-reduce")
    ap.add_argument("--no-graph", action="store_true",
                    help="run the step eagerly (default: the whole step -- forward, backward, clip, Adam -- is\n"
                         "captured in ONE CUDA graph on a single GPU; under DDP the step stays eager)")
    args = ap.parse_args()

    from inverse_flow_b200 import glow
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # stdout stays ONE JSON line: NCCL's log (whatever NCCL_DEBUG the caller set) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)

    torch.manual_seed(0)                       # same initial weights on every rank
    model, shape, batch = glow.build(args.model, args.coupling_width, args.groups or None, fused=args.fused)
    if args.init == "damped":
        with torch.no_grad():
            for inv in model.inv_layers:
                centre = inv.weight_fwd[:, :, -1, -1].clone()
                inv.weight_fwd.mul_(0.05)
                inv.weight_fwd[:, :, -1, -1] = centre
    model = model.to(device)
    init_gen = torch.Generator(device=device).manual_seed(99)                 # same batch on every rank
    model.initialize(torch.rand((batch, *shape), generator=init_gen, device=device) - 0.5)
    # Data parallelism without a wrapper: every parameter's .grad is a view into ONE flat bucket, summed over the
    # ranks by a single NCCL all-reduce per step -- which, unlike DistributedDataParallel's hooks, is captured in the
    # step's CUDA graph together with everything else (--ddp keeps the PyTorch wrapper, eager)
    use_ddp = world > 1 and args.ddp
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if use_ddp else model
    use_graph = not use_ddp and not args.no_graph
    params = [p for p in model.parameters() if p.requires_grad]
    flat_grad = torch.zeros(sum(p.numel() for p in params), device=device)
    off = 0
    for p in params:
        p.grad = flat_grad[off:off + p.numel()].view_as(p)
        off += p.numel()
    # the learning rate lives in a device tensor so that the warm-up also works inside a captured graph
    lr = torch.tensor(args.lr / max(args.lr_warmup_steps, 1), device=device)
    opt = torch.optim.Adam(model.parameters(), lr=lr, capturable=True)
    lr_step = torch.tensor(args.lr / max(args.lr_warmup_steps, 1), device=device)
    lr_max = torch.tensor(args.lr, device=device)
    gen = torch.Generator(device=device).manual_seed(1234 + rank)

    def batch_of_images():
        x = torch.randint(0, 256, (batch, *shape), generator=gen, device=device).float()
        x = (x + torch.rand(x.shape, generator=gen, device=device)) / 256.0 - 0.5        # dequantise, centre
        return x

    def train_step(x):
        opt.zero_grad(set_to_none=False)
        latents, logp = net(x)
        nll = -logp
        nll = torch.where(nll != nll, torch.zeros_like(nll), nll)       # NaN samples count as 0 (experiment.py:191)
        loss = nll.sum() / len(x) / (0.6931471805599453 * x[0].numel())
        loss.backward()
        if world > 1 and not use_ddp:
            dist.all_reduce(flat_grad)                                    # the one collective of the step (NCCL)
            flat_grad.mul_(1.0 / world)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        torch.minimum(lr + lr_step, lr_max, out=lr)                      # linear warm-up (experiment.py:197-202)
        return loss

    if use_graph:
        static_x = batch_of_images()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                      # warm-up outside capture (allocator, cuDNN, NCCL, our modules)
                train_step(static_x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = train_step(static_x)

        def step():
            static_x.copy_(batch_of_images())
            graph.replay()
            return static_loss
    else:
        def step():
            return train_step(batch_of_images())

    losses = []
    for _ in range(max(args.warmup, 3)):
        losses.append(float(step()))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    losses.append(float(loss))
    if rank == 0:
        n_inv = len(model.inv_layers)
        print(json.dumps({
            "metric": "if_glow train images/s", "unit": "images/s", "value": batch * world * args.steps / (ms * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
            "config": {"model": args.model, "input": list(shape), "batch_per_gpu": batch, "inv_conv_layers": n_inv,
                       "coupling_width": args.coupling_width, "groups": args.groups, "init": args.init, "fused_actnorm_squeeze": bool(args.fused),
                       "lr": args.lr, "lr_warmup_steps": args.lr_warmup_steps,
                       "parameters": sum(p.numel() for p in model.parameters()),
                       "parallelism": ("DistributedDataParallel over NCCL, one process per GPU" if use_ddp else
                                       "one process per GPU, one NCCL all-reduce of the flat gradient bucket per step "
                                       "(inside the step's CUDA graph)") if world > 1 else "single GPU",
                       "cuda_graph": bool(use_graph),
                       "note": "PyTorch layers around the inverse-conv kernels are minimal stand-ins (out of scope): a "
                               "drop-in / scaling check, not a tuned number"},
            "loss_bits_per_dim_first_last": [losses[0], losses[-1]],
        }))
    if world > 1:
        # a captured NCCL all-reduce keeps the communicator busy at tear-down (destroy_process_group then waits
        # for ever): synchronise, meet at a barrier, and leave without the destructor
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
