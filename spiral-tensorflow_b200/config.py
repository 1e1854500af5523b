"""Command-line flags of the reference (config.py:14-101) -- same names and defaults, so an
``args`` namespace built here can be handed to ``create_env`` / ``Policy`` / ``Discriminator``
exactly as in the reference.  Added flags are marked NEW."""
import argparse
from pathlib import Path


def str2bool(v):
    # reference utils/args.py:3-4: substring test against 'true'
    return v.lower() in ('true')


def int_list(v):
    return [int(x) for x in v.split(',')]


def build_parser():
    parser = argparse.ArgumentParser()
    # model
    parser.add_argument('--lstm_size', default=256, type=int)
    parser.add_argument('--scale', default=1.0, type=float)
    parser.add_argument('--z_dim', default=10, type=int)
    parser.add_argument('--dynamic_channel', default=False, type=str2bool)
    parser.add_argument('--disc_dim', default=64, type=int)
    parser.add_argument('--disc_batch_norm', default=True, type=str2bool)
    parser.add_argument('--loss', default='gan', type=str, choices=['l2', 'gan'])
    # environment
    parser.add_argument('--env', default="simple_mnist")
    parser.add_argument('--jump', default=True, type=str2bool)
    parser.add_argument('--curve', default=True, type=str2bool)
    parser.add_argument('--episode_length', default=5, type=int)
    parser.add_argument('--screen_size', default=64, type=int)
    parser.add_argument('--location_size', default=32, type=int)
    parser.add_argument('--color_channel', default=3, type=int, choices=[3, 1])
    parser.add_argument('--mnist_nums', default='0,1,2,3,4,5,6,7,8,9', type=int_list)
    parser.add_argument('--brush_path', default='assets/brushes/dry_brush.myb', type=str)
    parser.add_argument('--conditional', default=True, type=str2bool)
    # train
    parser.add_argument('--policy_lr', default=1e-5, type=float)
    parser.add_argument('--disc_lr', default=1e-4, type=float)
    parser.add_argument('--clip_disc_weights', default=False, type=str2bool)
    parser.add_argument('--entropy_coeff', default=0.01, type=float)
    parser.add_argument('--grad_clip', default=40, type=int)
    parser.add_argument('--policy_batch_size', default=64, type=int)
    parser.add_argument('--disc_batch_size', default=64, type=int)
    parser.add_argument('--replay_size', default=10, type=int)
    parser.add_argument('--buffer_batch_num', default=20, type=int)
    parser.add_argument('--wgan_lambda', default=20, type=float)
    parser.add_argument('--train', default=True, type=str2bool)
    # distributed (kept for CLI compatibility; the B200 build is synchronous data parallel)
    parser.add_argument('--task', default=0, type=int)
    parser.add_argument('--job_name', default="worker")
    parser.add_argument('--num_workers', default=4, type=int)
    parser.add_argument('--start_port', default=13333, type=int)
    parser.add_argument('--tag', default='spiral', type=str)
    # misc
    parser.add_argument('--debug', type=str2bool, default=False)
    parser.add_argument('--num_gpu', type=int, default=1)
    parser.add_argument('--policy_log_step', type=int, default=20)
    parser.add_argument('--disc_log_step', type=int, default=50)
    parser.add_argument('--data_dir', type=Path, default='.data')
    parser.add_argument('--log_dir', type=Path, default='logs')
    parser.add_argument('--load_path', type=Path, default=None)
    parser.add_argument('--log_level', type=str, default='INFO', choices=['INFO', 'DEBUG', 'WARN'])
    parser.add_argument('--seed', type=int, default=123)
    parser.add_argument('--dry_run', action='store_true')
    parser.add_argument('--tb_port', type=int, default=12345)
    # NEW: canvases per GPU, pending-dab capacity, read-back dither, D arithmetic
    parser.add_argument('--num_envs', type=int, default=64)
    parser.add_argument('--max_dabs', type=int, default=2560)
    parser.add_argument('--dither', type=str2bool, default=True)
    # arithmetic of the brush state machine (include/spiral_b200.h SpiralEnvCfg.math_mode): 'det' = double-precision,
    # correctly rounded exp / log / hypot (canvases bit-identical to the oracle's det mode, which matches glibc on >= 99 %
    # of canvases); 'fast' = binary32 polynomials and a division-light dab loop (bit-identical to the oracle's fast mode;
    # against libm: mean |err| <= 1e-3 on every canvas, single-dab jitter differences on ~20 % of 20-step canvases; the
    # default, DESIGN.md section 2 says why)
    parser.add_argument('--raster_math', type=str, default='fast', choices=['det', 'fast'])
    # trajectory store: RGB observations are integers 0..255, kept as uint8 (lossless, 4x smaller); False = float32
    parser.add_argument('--traj_uint8', type=str2bool, default=True)
    # several ranks: score the GAN reward with batch-norm statistics of the GLOBAL batch (per-layer all-reduce of the
    # channel sums, spiral_disc_forward_staged) instead of each rank's shard (the reference normalises with whatever
    # batch the learner dequeued, discriminator.py:76-79)
    parser.add_argument('--disc_global_bn', type=str2bool, default=False)
    parser.add_argument('--disc_precision', type=str, default='bf16x3', choices=['fp32', 'bf16x3', 'bf16'])
    # arithmetic of the WGAN-GP step's convolutions and their gradients (K2b); 'auto' = bf16x6 (24 operand bits,
    # fp32-exact products on the tensor cores) when the reward forward runs on tensor cores, else fp32
    # PyTorch policy trunk (the next row): TF32 convolutions / matmuls (4x faster on B200, ~1e-3 relative);
    # False = strict fp32 like the reference's TF-1.6 cuDNN path
    parser.add_argument('--policy_tf32', type=str2bool, default=True)
    # acting path (policy.act in the rollout): trunk convolutions through libspiral_b200.so in bf16 (the learner stays fp32)
    parser.add_argument('--policy_act_native', type=str2bool, default=True)
    # learner path (policy forward / backward over the [B*T] frames): convolutions and their gradients through
    # libspiral_b200.so (csrc/policy_learn.cu, bf16x3: fp32 parity); False = the PyTorch module's cuDNN convolutions
    parser.add_argument('--policy_learn_native', type=str2bool, default=True)
    # operand split of those kernels: 'bf16x6' = three bf16 planes / six products, fp32-exact operands (logits and gradients
    # as close to float64 as PyTorch's strict-fp32 path); 'bf16x3' = two planes / three products, 16 operand bits (~2x
    # faster, per-layer error ~1e-5: 64x finer than TF32)
    parser.add_argument('--policy_learn_precision', type=str, default='f16x3', choices=['f16x3', 'bf16x6', 'bf16x3'])
    # capture the T-step rollout (policy.act + env.step) into one CUDA graph and replay it
    parser.add_argument('--cuda_graphs', type=str2bool, default=True)
    parser.add_argument('--disc_grad_precision', type=str, default='auto',
                        choices=['auto', 'fp32', 'bf16x3', 'bf16', 'bf16x6'])
    return parser


def get_args(argv=None, parse_unknown=False, check_workers=False):
    """reference config.py:83-101, including the override of ``conditional`` by ``loss``."""
    parser = build_parser()
    if parse_unknown:
        args, unknown = parser.parse_known_args(argv)
    else:
        args = parser.parse_args(argv)
    if args.loss == 'gan':
        args.conditional = False
        if check_workers:
            assert args.num_workers > 2, "num_workers should be larger than 2 (policy, discriminator, worker)"
    elif args.loss == 'l2':
        args.conditional = True
        if check_workers:
            assert args.num_workers > 1, "num_workers should be larger than 2 (policy, worker)"
    if parse_unknown:
        return args, unknown
    return args
