#!/usr/bin/env python
"""Config-driven training entry point -- the reference's ``train.py`` contract on the B200 path.

    python train.py --config_file config/pds-coco/zeng-bihome-lr-1e-3.yaml
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 train.py --config_file ...

Same YAML files, same ``main(config_file_path)`` / ``do_train`` / ``train_one_epoch`` / ``eval_one_epoch`` split,
same TensorBoard scalar names, same checkpoint files as reference ``train.py:284-757``.  What is different:

* the input pipeline is the GPU pair generator (K5, ``bihome_b200.data.gpu_pairs``) over a uint8 image pool resident
  in HBM instead of 8 DataLoader worker processes (reference ``train.py:80-137``); the pool is read from
  ``DATA.TRAIN_SPLIT`` (``.npy`` files of ``preprocess_offline.py``, or ``.jpg`` rescaled the same way) and falls
  back to a synthetic pool when that directory does not exist;
* multi-GPU is one process per GPU under ``torchrun`` with NCCL gradient all-reduce (DDP); the reference's
  ``nn.DataParallel`` branch (``train.py:513-518``) cannot run a scalar-loss head.  ``SAMPLER.BATCH_SIZE`` is per
  GPU, every rank draws a disjoint stream (``gpu_pairs.rank_seed``), rank 0 logs and checkpoints;
* nothing in the step synchronises with the host except on logging steps.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bihome_b200 import engine  # noqa: E402
from bihome_b200 import functional as F  # noqa: E402
from bihome_b200.data import gpu_pairs  # noqa: E402
from bihome_b200.utils.checkpoint import CheckPointer  # noqa: E402

STRING_LOSSES = ('TripletLoss', 'iHomE', 'biHomE')


class _NullWriter:
    """stands in for SummaryWriter on ranks > 0 and when tensorboard is not installed"""

    def add_scalars(self, *a, **k):
        pass

    def flush(self):
        pass

    def close(self):
        pass


def make_summary_writer(log_dir, enabled=True):
    if not enabled:
        return _NullWriter()
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(log_dir)
    except Exception as e:  # noqa: BLE001
        print('train.py: tensorboard unavailable (%s); scalars are not written' % e)
        return _NullWriter()


def dist_env():
    """(rank, local_rank, world) from the torchrun environment; (0, 0, 1) for a plain `python train.py`"""
    return int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


make_pair_loader = gpu_pairs.loader_from_config      # reference make_coco_dataloader, train.py:80-137


def forward_loss(model, data, loss_fn):
    """the reference's loss dispatch (train.py:319-352): returns (loss, delta_gt, delta_hat)"""
    if isinstance(loss_fn, torch.nn.Module):
        ground_truth, network_output, delta_gt, delta_hat = model(data)
        return loss_fn(ground_truth, network_output), delta_gt, delta_hat
    if loss_fn == 'CosineDistance':
        ground_truth, network_output, delta_gt, delta_hat = model(data)
        return torch.sum(1 - torch.cosine_similarity(ground_truth, network_output, dim=1)), delta_gt, delta_hat
    if loss_fn in STRING_LOSSES:
        return model(data)
    raise AssertionError('Do not know the loss: ' + str(loss_fn))


def train_one_epoch(model, loader, optimizer, gradient_clip, scheduler, loss_fn, epoch, steps_per_epoch, checkpointer,
                    checkpoint_arguments, log_step, summary_writer, self_supervised=False, log_verbose=False, max_steps=None,
                    cuda_graph=False):
    model.train()
    graphed = None
    # a checkpoint written in the middle of an epoch (--max_steps) resumes at the iteration after it: no step is repeated
    # and no TensorBoard step number is reused
    skip = min(max(checkpoint_arguments['step'] - epoch * steps_per_epoch, 0), steps_per_epoch)
    step = epoch * steps_per_epoch + skip
    batches = loader.batches(steps_per_epoch - skip) if hasattr(loader, 'batches') else loader
    for iter_no, data in enumerate(batches, start=skip):
        step = epoch * steps_per_epoch + iter_no + 1
        logging_step = step % log_step == 0
        if cuda_graph and loss_fn in STRING_LOSSES and not logging_step:
            # forward + backward replayed from one CUDA graph (engine.GraphedStep); logging steps stay eager (they read
            # scalars back to the host)
            if graphed is None:
                graphed = engine.GraphedStep(model, data)
            loss, delta_gt, delta_hat = engine.graphed_train_step(graphed, data, optimizer, scheduler, gradient_clip)
            if max_steps is not None and step >= max_steps:
                break
            continue
        if graphed is None:
            optimizer.zero_grad(set_to_none=True)
        else:
            optimizer.zero_grad(set_to_none=False)      # the graph owns the gradient buffers
        if logging_step and not isinstance(summary_writer, _NullWriter):
            data['summary_writer'] = summary_writer
            data['summary_writer_step'] = step
        loss, delta_gt, delta_hat = forward_loss(model, data, loss_fn)
        loss.backward()
        if cuda_graph and not isinstance(model, torch.nn.parallel.DistributedDataParallel):
            # --cuda_graph under torchrun: no DDP wrapper (engine.GraphedStep); eager steps exchange their gradients the same way
            engine.average_gradients([p for p in model.parameters() if p.requires_grad], graphed.flat if graphed is not None else None)
        if gradient_clip > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), gradient_clip)
        optimizer.step()
        scheduler.step()
        if logging_step:
            grads = [p.grad for p in model.parameters() if p.grad is not None]
            total_norm = float(torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads])))
            if self_supervised and delta_gt is not None:
                summary_writer.add_scalars('mace', {'train': float(F.mace(delta_gt.float(), delta_hat.float()))}, step)
            summary_writer.add_scalars('loss', {'train': loss.item()}, step)
            summary_writer.add_scalars('lr', {'value': scheduler.get_last_lr()[0]}, step)
            summary_writer.add_scalars('g_norm', {'value': total_norm}, step)
            summary_writer.flush()
            if log_verbose:
                print('Epoch: {} iter: {}/{} loss: {}'.format(epoch, iter_no + 1, steps_per_epoch, loss.item()))
        if max_steps is not None and step >= max_steps:
            break
    checkpoint_arguments['step'] = step
    checkpointer.save('model_{:06d}'.format(step), **checkpoint_arguments)
    return step


def eval_one_epoch(model, loader, loss_fn, epoch, steps_per_epoch, summary_writer, self_supervised=False, log_verbose=False,
                   rank=0, world=1):
    """reference train.py:432-487.  Under DDP the test batches are dealt round-robin to the ranks (batch i of the seeded
    test stream goes to rank i % world) and the three sums meet in one all-reduce: no rank idles in a collective while
    rank 0 evaluates alone, and the mean is the one a single process computes."""
    model.eval()
    losses, maces = [], []
    with torch.no_grad():
        shard = loader.shard(rank, world) if world > 1 else loader
        for iter_no, data in enumerate(shard):
            loss, delta_gt, delta_hat = forward_loss(model, data, loss_fn)
            losses.append(loss.detach().float())
            if self_supervised and delta_gt is not None:
                maces.append(F.mace(delta_gt.float(), delta_hat.float()))
            if log_verbose:
                print('Epoch: {} iter: {}/{} loss: {}'.format(epoch, iter_no + 1, len(loader), loss.item()))
    dev = next(model.parameters()).device
    zero = torch.zeros((), device=dev)
    sums = torch.stack([torch.stack(losses).sum() if losses else zero, zero + len(losses),
                        torch.stack(maces).sum() if maces else zero, zero + len(maces)]).double()
    if world > 1:
        dist.all_reduce(sums)
    sums = sums.tolist()
    mean_loss = sums[0] / sums[1] if sums[1] else float('nan')
    mean_mace = sums[2] / sums[3] if sums[3] else float('nan')
    maces = sums[3] > 0
    summary_writer.add_scalars('loss', {'test': mean_loss}, (epoch + 1) * steps_per_epoch)
    if maces:
        summary_writer.add_scalars('mace', {'test': mean_mace}, (epoch + 1) * steps_per_epoch)
    summary_writer.flush()
    return mean_loss, mean_mace


def do_train(model, train_loader, test_loader, optimizer, gradient_clip, scheduler, loss_fn, epochs, steps_per_epoch,
             checkpointer, checkpoint_arguments, log_dir='logs', log_step=1, self_supervised=False, log_verbose=False,
             rank=0, max_steps=None, world=1, cuda_graph=False):
    writer = make_summary_writer(log_dir, enabled=rank == 0)
    start_epoch = checkpoint_arguments['step'] // steps_per_epoch
    if hasattr(train_loader, 'step'):
        train_loader.step = checkpoint_arguments['step']      # the pair stream continues where the checkpoint left it
    for epoch in range(start_epoch, epochs):
        if rank == 0:
            print('Training epoch: {}'.format(epoch))
        t0 = time.perf_counter()
        step = train_one_epoch(model, train_loader, optimizer, gradient_clip, scheduler, loss_fn, epoch, steps_per_epoch,
                               checkpointer, checkpoint_arguments, log_step, writer, self_supervised, log_verbose, max_steps,
                               cuda_graph=cuda_graph)
        torch.cuda.synchronize()
        if rank == 0:
            done = step - epoch * steps_per_epoch
            print('  {} steps in {:.1f} s ({:.0f} image pairs/s per GPU)'.format(
                done, time.perf_counter() - t0, done * train_loader.batch_size / (time.perf_counter() - t0)))
        if test_loader is not None:
            if rank == 0:
                print('Testing epoch: {}'.format(epoch))
            test_loader.step = 0        # the same test pairs every epoch (reference: seeded sampler + seeded transforms)
            mean_loss, mean_mace = eval_one_epoch(_plain(model), test_loader, loss_fn, epoch, steps_per_epoch, writer,
                                                  self_supervised, log_verbose, rank, world)
            if rank == 0:
                print('  test loss {:.4f}  MACE {:.4f}'.format(mean_loss, mean_mace))
        if max_steps is not None and step >= max_steps:
            break
    writer.close()


def _plain(model):
    return model.module if isinstance(model, torch.nn.parallel.DistributedDataParallel) else model


def main(config_file_path, batch_size=None, max_steps=None, synthetic_pool=256, channels_last=True, log_dir=None, init_seed=0,
         cuda_graph=False, cudnn_benchmark=True):
    config = engine.load_config(config_file_path)
    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit('train.py: no CUDA device -- the biHomE hot path runs on sm_100a kernels only (no CPU fallback)')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    # fixed shapes for the whole run: let cuDNN time its convolution algorithms once per shape (3.8 % of the step on a B200)
    torch.backends.cudnn.benchmark = cudnn_benchmark
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=device)

    if 'coco' not in config['DATA']['NAME']:
        raise NotImplementedError('the GPU pair source replaces the coco pipeline only (DATA.NAME = %r)' % config['DATA']['NAME'])
    train_loader = make_pair_loader(config, 'train', device, rank, batch_size, synthetic_pool)
    test_loader = make_pair_loader(config, 'test', device, 0, batch_size, max(8, synthetic_pool // 8)) \
        if 'TEST_SPLIT' in config['DATA'] else None

    torch.manual_seed(init_seed)    # identical initial weights on every rank (--init_seed; the reference leaves it unseeded)
    model = engine.build_model(config).to(device)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    solver = config['SOLVER']
    optimizer, scheduler = engine.build_optimizer(config, model)
    try:
        loss_fn = getattr(torch.nn, solver['LOSS'])()
    except AttributeError:
        loss_fn = solver['LOSS']    # 'biHomE', 'iHomE', 'TripletLoss', 'CosineDistance' (reference train.py:712-715)
    gradient_clip = solver['GRADIENT_CLIP'] if 'GRADIENT_CLIP' in solver else -1

    log_dir = log_dir or config['LOGGING']['DIR']
    restart_lr = 'RESTART_LEARNING_RATE' in solver and solver['RESTART_LEARNING_RATE']
    # reference train.py:723-728: RESTART_LEARNING_RATE withholds the optimizer only; the scheduler is always restored, so
    # the milestones stay aligned with the global step
    checkpointer = CheckPointer(model, None if restart_lr else optimizer, scheduler, log_dir,
                                save_to_disk=rank == 0, device=str(device))
    extra = checkpointer.load()
    arguments = {'step': 0}
    arguments.update(extra)
    if 'PRETRAINED' in config['MODEL'] and arguments['step'] == 0:
        blob = torch.load(config['MODEL']['PRETRAINED'], map_location='cpu', weights_only=False)
        model.load_state_dict(blob['model'])
    if restart_lr:
        checkpointer.optimizer = optimizer

    net = model
    if world > 1 and cuda_graph and isinstance(loss_fn, str):
        # graph replay per rank + one flat all-reduce of the gradients (engine.GraphedStep): no DDP wrapper
        engine.sync_module_state(model)
    elif world > 1:
        net = engine.data_parallel(model, local)
        checkpointer.model = net
    self_supervised = 'SELF_SUPERVISED' in config['DATA'] and config['DATA']['SELF_SUPERVISED'] or isinstance(loss_fn, str)
    do_train(net, train_loader, test_loader, optimizer, gradient_clip, scheduler, loss_fn, solver['NUM_EPOCHS'],
             len(train_loader), checkpointer, arguments, log_dir=log_dir, log_step=config['LOGGING']['STEP'],
             self_supervised=self_supervised, log_verbose=config['LOGGING'].get('VERBOSE', False), rank=rank, max_steps=max_steps,
             world=world, cuda_graph=cuda_graph)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--config_file', type=str, required=True, help='path to the YAML config')
    ap.add_argument('--batch_size', type=int, default=None, help='override DATA.SAMPLER.BATCH_SIZE (per GPU)')
    ap.add_argument('--max_steps', type=int, default=None, help='stop early (smoke runs)')
    ap.add_argument('--synthetic_pool', type=int, default=256, help='synthetic images per GPU when the dataset is absent')
    ap.add_argument('--log_dir', type=str, default=None, help='override LOGGING.DIR')
    ap.add_argument('--nchw', dest='channels_last', action='store_false')
    ap.add_argument('--init_seed', type=int, default=0, help='torch seed for the initial weights (same on every rank)')
    ap.add_argument('--cuda_graph', action='store_true',
                    help='replay forward + backward from one CUDA graph (single GPU; pays off at small, launch-bound batches)')
    ap.add_argument('--no_cudnn_benchmark', dest='cudnn_benchmark', action='store_false')
    a = ap.parse_args()
    main(a.config_file, a.batch_size, a.max_steps, a.synthetic_pool, a.channels_last, a.log_dir, a.init_seed, a.cuda_graph,
         a.cudnn_benchmark)
