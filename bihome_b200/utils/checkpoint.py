"""Checkpoint save / resume in the reference's on-disk format (``src/utils/checkpoint.py:10-115``):

    <LOGGING.DIR>/model_%06d.pth   = {'model': state_dict, 'optimizer': ..., 'scheduler': ..., 'step': int}
    <LOGGING.DIR>/last_checkpoint.txt  -> path of the newest file

``model`` is the ``nn.Sequential(backbone, head)`` state dict (keys ``0.*``, ``1.backbone.*``,
``1.auxiliary_resnet.resnet.*``), so files written by the reference load here and vice versa.  Under DDP only
rank 0 writes (``save_to_disk``), and the wrapper is unwrapped before ``state_dict()``.
"""
import logging
import os

import torch

LAST = 'last_checkpoint.txt'


def _unwrap(model):
    return model.module if isinstance(model, torch.nn.parallel.DistributedDataParallel) else model


class CheckPointer:

    def __init__(self, model, optimizer=None, scheduler=None, save_dir='', save_to_disk=None, logger=None, device=None):
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.save_dir, self.save_to_disk = save_dir, save_to_disk
        self.logger = logger or logging.getLogger(__name__)
        self.device = device

    # ---- writing ---------------------------------------------------------------------------------
    def save(self, name, **extra):
        if not self.save_dir or not self.save_to_disk:
            return None
        blob = {'model': _unwrap(self.model).state_dict()}
        if self.optimizer is not None:
            blob['optimizer'] = self.optimizer.state_dict()
        if self.scheduler is not None:
            blob['scheduler'] = self.scheduler.state_dict()
        blob.update(extra)
        os.makedirs(self.save_dir, exist_ok=True)
        path = os.path.join(self.save_dir, '{}.pth'.format(name))
        self.logger.info('Saving checkpoint to %s', path)
        torch.save(blob, path)
        self.tag_last_checkpoint(path)
        return path

    def tag_last_checkpoint(self, filename):
        with open(os.path.join(self.save_dir, LAST), 'w') as f:
            f.write(filename)

    # ---- reading ---------------------------------------------------------------------------------
    def has_checkpoint(self):
        return bool(self.save_dir) and os.path.exists(os.path.join(self.save_dir, LAST))

    def get_checkpoint_file(self):
        try:
            with open(os.path.join(self.save_dir, LAST)) as f:
                return f.read().strip()
        except IOError:
            return ''

    def load(self, f=None, use_latest=True):
        """restore model (+ optimizer / scheduler when given); returns what else the file holds (e.g. {'step': n})"""
        if f is None and use_latest and self.has_checkpoint():
            f = self.get_checkpoint_file()
        if not f:
            self.logger.info('No checkpoint found.')
            return {}
        self.logger.info('Loading checkpoint from %s', f)
        blob = torch.load(f, map_location='cpu', weights_only=False)
        _unwrap(self.model).load_state_dict(blob.pop('model'))
        if 'optimizer' in blob and self.optimizer is not None:
            self.optimizer.load_state_dict(blob.pop('optimizer'))   # torch moves the state to the parameters' device
        if 'scheduler' in blob and self.scheduler is not None:
            sched = blob.pop('scheduler')
            # like the reference: only the counters are restored, milestones come from the config
            self.scheduler._step_count = sched['_step_count']
            self.scheduler.last_epoch = sched['last_epoch']
        return blob
