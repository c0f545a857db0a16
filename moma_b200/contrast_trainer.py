"""Contrast trainer -- host-side mirror of the reference learning/contrast_trainer.py
and learning/base_trainer.py (the parts helper/loops_moma.py:308-335 and
train_student_moma.py:229-336 call).

On the B200 path:
  * ``momentum_update``  -> one multi-tensor EMA launch (reference :207-211, 2 launches/tensor);
  * ``_global_gather``   -> ``all_gather_into_tensor`` straight into the [W*B, D] result
                            (reference :83-88: W ``ones_like`` allocations + ``cat``);
  * ``_compute_loss_accuracy`` -> served by the LazyLogits handle, no [B, K+1] passes, no host sync.
``_shuffle_bn`` exchanges image SLICES (one all-to-all) instead of all-gathering the node's images;
its permutation, its returned (k, all_k) and the batches the momentum encoder sees are the reference's.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .lazy_logits import LazyLogits


_PEER_GATHER_MAX_BYTES = int(os.environ.get("MOMA_B200_PEER_GATHER_MAX_BYTES", 64 << 20))


class AverageMeter(object):
    """Computes and stores the average and current value (reference learning/util.py:7-22)"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def accuracy(output, target, topk=(1,)):
    """Top-k accuracy (reference learning/util.py:25-41).  For a LazyLogits handle and k == 1
    the answer comes from the fused kernel's argmax-is-positive flags (labels are all zero)."""
    with torch.no_grad():
        batch_size = target.size(0)
        if isinstance(output, LazyLogits) and tuple(topk) == (1,) and output._targets_are_zero(target):
            return [output.top1_accuracy]
        maxk = max(topk)
        _, pred = output.topk(maxk, 1, True, True)
        pred = pred.t()
        correct = pred.eq(target.view(1, -1).expand_as(pred))
        res = []
        for k in topk:
            correct_k = correct[:k].reshape(-1).float().sum(0, keepdim=True)
            res.append(correct_k.mul_(100.0 / batch_size))
        return res


class BaseTrainer(object):
    """reference learning/base_trainer.py:15-91"""

    def __init__(self, args):
        self.args = args
        self.local_group = None
        self.logger = None

    def init_ddp_environment(self, gpu, ngpus_per_node):
        a = self.args
        a.gpu = gpu
        a.ngpus_per_node = ngpus_per_node
        a.node_rank = a.rank
        a.local_rank = gpu
        a.local_center = a.rank * ngpus_per_node
        if torch.cuda.is_available():
            torch.cuda.set_device(gpu)
            torch.backends.cudnn.benchmark = True
        if a.gpu is not None:
            print("Use GPU: {} for training".format(a.gpu))
        if a.multiprocessing_distributed:
            a.rank = a.rank * ngpus_per_node + gpu
            os.environ['PYTHONWARNINGS'] = 'ignore:semaphore_tracker:UserWarning'
            dist.init_process_group(backend=a.dist_backend, init_method=a.dist_url,
                                    world_size=a.world_size, rank=a.rank)
        # per-node group for ShuffleBN
        local_groups = []
        for i in range(0, a.world_size // ngpus_per_node):
            gp = dist.new_group(ranks=list(range(i * ngpus_per_node, (i + 1) * ngpus_per_node)),
                                backend=a.dist_backend)
            local_groups.append(gp)
        self.local_group = local_groups[a.rank // ngpus_per_node]
        if a.local_rank == 0:
            print("node_rank:", a.node_rank)
            print("local_center:", a.local_center)
            print("local group size:", dist.get_world_size(self.local_group))

    def init_tensorboard_logger(self):
        if self.args.rank == 0:
            import tensorboard_logger as tb_logger
            self.logger = tb_logger.Logger(logdir=self.args.tb_folder, flush_secs=2)

    def adjust_learning_rate(self, optimizer, epoch):
        args = self.args
        lr = args.learning_rate
        if args.cosine:
            eta_min = lr * (args.lr_decay_rate ** 3)
            lr = eta_min + (lr - eta_min) * (1 + math.cos(math.pi * epoch / args.epochs)) / 2
        else:
            steps = np.sum(epoch > np.asarray(args.lr_decay_epochs))
            if steps > 0:
                lr = lr * (args.lr_decay_rate ** steps)
        for group in optimizer.param_groups:
            group['lr'] = lr

    def warmup_learning_rate(self, epoch, batch_id, total_batches, optimizer):
        args = self.args
        if args.warm and epoch <= args.warm_epochs:
            p = (batch_id + (epoch - 1) * total_batches) / (args.warm_epochs * total_batches)
            lr = args.warmup_from + p * (args.warmup_to - args.warmup_from)
            for group in optimizer.param_groups:
                group['lr'] = lr


class ContrastTrainer(BaseTrainer):
    """trainer for contrastive pretraining (reference learning/contrast_trainer.py:19-211)"""

    def __init__(self, args):
        super().__init__(args)

    def logging(self, epoch, logs, lr):
        if self.args.rank == 0:
            for name, v in zip(('loss', 'acc', 'jig_loss', 'jig_acc'), logs):
                self.logger.log_value(name, v, epoch)
            self.logger.log_value('learning_rate', lr, epoch)

    def wrap_up(self, model, model_ema, optimizer):
        """DDP wrap (+ apex amp when args.amp); unused by the moma driver (reference :40-69)."""
        from torch.nn.parallel import DistributedDataParallel as DDP
        args = self.args
        model.cuda(args.gpu)
        if isinstance(model_ema, torch.nn.Module):
            model_ema.cuda(args.gpu)
        if getattr(args, "amp", False):
            from apex import amp
            model, optimizer = amp.initialize(model, optimizer, opt_level=args.opt_level)
            if isinstance(model_ema, torch.nn.Module):
                model_ema = amp.initialize(model_ema, opt_level=args.opt_level)
        model = DDP(model, device_ids=[args.gpu])
        if isinstance(model_ema, torch.nn.Module):
            self.momentum_update(model.module, model_ema, 0)
        return model, model_ema, optimizer

    def broadcast_memory(self, contrast):
        """Synchronise the memory buffers from rank 0 (reference :71-81)."""
        if getattr(contrast, "is_sharded", False):
            contrast.broadcast_from_rank0()
            return
        if self.args.mem in ['MoCo', 'MoCoAtt']:
            dist.broadcast(contrast.memory, 0)
        else:
            dist.broadcast(contrast.memory_s, 0)
            dist.broadcast(contrast.memory_t, 0)
        if hasattr(contrast, "invalidate_shadows"):      # the collective wrote the masters behind autograd's back
            contrast.invalidate_shadows()

    @staticmethod
    def _global_gather(x):
        """all_gather + cat(dim=0) -> [W*B, D]  (reference :83-88)"""
        world = dist.get_world_size()
        x = x.contiguous()
        # Gathers above MOMA_B200_PEER_GATHER_MAX_BYTES take the NCCL collective instead of the peer-memory kernel (whose
        # in-band tags double the bytes).  Measured at 8 GPUs on the C3 step (profiles/r02_scaling.txt): NCCL for the
        # 6.3 MB projection gather was NOT faster end to end (0.331 vs 0.291 ms per step), so the default keeps every
        # gather that fits the symmetric buffer on the peer kernel; what matters is that large and small exchanges do not
        # overlap in time (step.py sequences them).
        small = world * x.numel() * x.element_size() <= _PEER_GATHER_MAX_BYTES
        if small and x.is_cuda and x.dtype == torch.float32 and (x.numel() * 4) % 16 == 0:
            from .peer import CH_KEYS, PeerExchange
            peer = PeerExchange.create(None, x.device, 2 * world * x.numel() * 4)
            if peer is not None:                       # one NVLink push/flag/wait kernel instead of the collective
                return peer.allgather(x, CH_KEYS)
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x)
        return out

    # ---- ShuffleBN (reference :90-133) -----------------------------------------------------------------------
    # The reference all-gathers the INPUT IMAGES of the node (W x B x 3 x H x W: 38.5 MB per rank and step at
    # 64 x 3 x 224^2, 201 MB at 512^2) and then keeps the B rows its slice of the shared permutation names.  A rank only
    # needs those B images: with the permutation known everywhere, each rank sends every peer exactly the rows that
    # peer will feed to its momentum encoder (one all-to-all of image slices, (W-1)/W of B images per rank instead of
    # (W-1) x B received), and the result is bit-identical to ``node_x[this_ids]``.  The permutation itself is drawn
    # exactly as the reference draws it (torch.randperm on the host RNG, rank 0's draw broadcast), so the RNG streams and
    # the batches the momentum encoder's train-mode BatchNorm sees are the reference's.
    # MOMA_B200_SHUFFLE_BN=gather restores the reference's image all-gather (A/B switch).
    @staticmethod
    def _exchange_shuffled_rows(x, shuffle_ids_host, rank, world, bsz, group):
        """rows ``node_x[shuffle_ids[rank * bsz:(rank + 1) * bsz]]`` of the (never materialised) node batch
        ``node_x = cat(x of every rank)``, via one all-to-all of row slices."""
        ids = shuffle_ids_host.view(world, bsz)                  # ids[r] = the node rows rank r consumes, in its order
        owner = ids // bsz                                       # which rank holds each of them
        send_idx, in_splits = [], []
        for r in range(world):                                   # what I send to r: my rows among ids[r], in r's order
            mine = ids[r][owner[r] == rank] % bsz
            send_idx.append(mine)
            in_splits.append(int(mine.numel()))
        out_splits = [int((owner[rank] == s).sum()) for s in range(world)]
        send = x.index_select(0, torch.cat(send_idx).to(x.device))
        recv = torch.empty((bsz,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_to_all_single(recv, send, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
        # recv is grouped by source rank; put every row at its position in this rank's order
        pos = torch.cat([torch.nonzero(owner[rank] == s).flatten() for s in range(world)]).to(x.device)
        out = torch.empty_like(recv)
        out.index_copy_(0, pos, recv)
        return out

    def _shuffle_bn(self, x, model_ema, model_ema_head):
        """Shuffle-BN teacher forward (reference :90-133): returns (k, all_k)."""
        args = self.args
        local_gp = self.local_group
        bsz = x.size(0)
        wl = dist.get_world_size(local_gp)
        x = x.contiguous()
        sliced = wl > 1 and os.environ.get("MOMA_B200_SHUFFLE_BN", "slices") != "gather"

        shuffle_ids = torch.randperm(bsz * wl).to(x.device)      # same host-RNG draw as the reference (:108-110)
        reverse_ids = torch.argsort(shuffle_ids)
        dist.broadcast(shuffle_ids, 0)
        dist.broadcast(reverse_ids, 0)

        with torch.no_grad():
            if sliced:
                this_x = self._exchange_shuffled_rows(x, shuffle_ids.cpu(), args.local_rank, wl, bsz, local_gp)
            else:
                node_x = torch.empty((wl * bsz,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
                dist.all_gather_into_tensor(node_x, x, group=local_gp)
                this_x = node_x[shuffle_ids[args.local_rank * bsz:(args.local_rank + 1) * bsz]]
            feat_t, logit_t = model_ema(this_x, is_feat=True)
            k = model_ema_head(feat_t[-1])

        all_k = self._global_gather(k)

        node_id, ngpus = args.node_rank, args.ngpus_per_node
        node_k = all_k[node_id * ngpus * bsz:(node_id + 1) * ngpus * bsz]
        this_ids = reverse_ids[args.local_rank * bsz:(args.local_rank + 1) * bsz]
        return node_k[this_ids], all_k

    def _shuffle_bn_attn(self, x, model_ema, model_ema_head, criterion_kd, q):
        """Variant applying the attention before the gather (reference :135-187)."""
        args = self.args
        local_gp = self.local_group
        bsz = x.size(0)
        wl = dist.get_world_size(local_gp)
        x = x.contiguous()
        node_x = torch.empty((wl * bsz,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(node_x, x, group=local_gp)
        shuffle_ids = torch.randperm(bsz * wl).to(x.device)
        reverse_ids = torch.argsort(shuffle_ids)
        dist.broadcast(shuffle_ids, 0)
        dist.broadcast(reverse_ids, 0)
        this_ids = shuffle_ids[args.local_rank * bsz:(args.local_rank + 1) * bsz]
        with torch.no_grad():
            feat_t, logit_t = model_ema(node_x[this_ids], is_feat=True)
            k = model_ema_head(feat_t[-1])
        if args.attn == 'self_mix':
            out = criterion_kd.atts(torch.cat([q, k], dim=0))
            q, k = out[:bsz], out[bsz:]
        else:
            q = criterion_kd.atts_q(q)
            k = criterion_kd.atts_k(k)
        all_k = self._global_gather(k)
        node_id, ngpus = args.node_rank, args.ngpus_per_node
        node_k = all_k[node_id * ngpus * bsz:(node_id + 1) * ngpus * bsz]
        this_ids = reverse_ids[args.local_rank * bsz:(args.local_rank + 1) * bsz]
        return q, node_k[this_ids], all_k

    @staticmethod
    def _compute_loss_accuracy(logits, target, criterion):
        """losses / top-1 accuracies for a list of logits (reference :189-205)."""
        losses = [criterion(logit, target) for logit in logits]
        accuracies = [accuracy(logit, target)[0] for logit in logits]
        return losses, accuracies

    @staticmethod
    def momentum_update(model, model_ema, m):
        """model_ema = m * model_ema + (1 - m) * model over zip(parameters)  (reference :207-211)"""
        ops.ema_update([p.detach() for p in model.parameters()],
                       [p.detach() for p in model_ema.parameters()], m)
