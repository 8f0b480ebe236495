"""Timestep samplers of the training loop — the ``t_sampler`` argument of ``train_step`` (train.py:157, 261) and the
loss-history statistics the loop keeps beside it (train.py:571-600, 625), with the call surface of
/root/reference/core/resample.py: ``create_named_schedule_sampler(name, T)``, ``sampler.sample(batch_size, device) ->
(timesteps int64 [B], weights fp32 [B])``, ``sampler.weights()``, ``update_with_all_losses(ts, losses)``,
``update_with_local_losses(local_ts, local_losses)``.

Host-side by nature (a handful of integers per step; the draw uses numpy's global generator exactly like the reference,
``np.random.choice(T, size=B, p=p)``, so a seeded run draws the same timesteps).  The loss history is a ring buffer per
timestep instead of a shifted array: the weights only use the mean of the squared entries, which does not depend on
their order.
"""
from __future__ import annotations

import numpy as np
import torch


class ScheduleSampler:
    """Importance sampling over the T training timesteps: ``weights()`` (positive, not necessarily normalised) defines
    the proposal; ``sample`` returns indices and the factors ``1 / (T p_t)`` that keep the loss expectation unchanged."""

    def weights(self) -> np.ndarray:
        raise NotImplementedError

    def sample(self, batch_size: int, device):
        w = self.weights()
        p = w / np.sum(w)
        drawn = np.random.choice(len(p), size=(batch_size,), p=p)
        scale = 1 / (len(p) * p[drawn])
        return (torch.from_numpy(drawn).long().to(device), torch.from_numpy(scale).float().to(device))


class UniformSampler(ScheduleSampler):
    def __init__(self, T: int):
        self.T = T
        self._weights = np.ones([T])

    def weights(self):
        return self._weights


class SNRSampler(ScheduleSampler):
    """Proposal proportional to a caller-supplied per-timestep array (train.py:571 passes ``sqrt(snr) + 1``)."""

    def __init__(self, snr):
        self._snr = snr

    def weights(self):
        return self._snr


class LossAwareSampler(ScheduleSampler):
    def update_with_all_losses(self, ts, losses):
        raise NotImplementedError

    def update_with_local_losses(self, local_ts: torch.Tensor, local_losses: torch.Tensor):
        """Every rank contributes its (timestep, loss) pairs; all ranks then apply the same update in rank order.  Batches
        may differ in length between ranks: sizes are gathered first and the payloads padded to the longest."""
        import torch.distributed as dist
        world = dist.get_world_size()
        n_local = torch.tensor([len(local_ts)], dtype=torch.int32, device=local_ts.device)
        sizes = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(sizes, n_local)
        sizes = [int(s.item()) for s in sizes]
        longest = max(sizes)

        def padded(x):
            out = torch.zeros(longest, dtype=x.dtype, device=x.device)
            out[:len(x)] = x
            return out

        ts_all = [torch.zeros(longest, dtype=local_ts.dtype, device=local_ts.device) for _ in range(world)]
        loss_all = [torch.zeros(longest, dtype=local_losses.dtype, device=local_losses.device) for _ in range(world)]
        dist.all_gather(ts_all, padded(local_ts))
        dist.all_gather(loss_all, padded(local_losses))
        ts = torch.cat([t[:n] for t, n in zip(ts_all, sizes)]).tolist()
        losses = torch.cat([v[:n] for v, n in zip(loss_all, sizes)]).tolist()
        self.update_with_all_losses(ts, losses)


class LossSecondMomentResampler(LossAwareSampler):
    """Proposal proportional to the root mean square of the last ``history_per_term`` losses seen at each timestep, mixed
    with ``uniform_prob`` of the uniform distribution; uniform until every timestep has a full history."""

    def __init__(self, T: int, history_per_term: int = 10, uniform_prob: float = 0.001):
        self.T = T
        self.history_per_term = history_per_term
        self.uniform_prob = uniform_prob
        self._loss_history = np.zeros([T, history_per_term], dtype=np.float64)   # ring buffer per timestep
        self._loss_counts = np.zeros([T], dtype=np.int32)                        # entries filled (saturates)
        self._head = np.zeros([T], dtype=np.int64)                               # next slot to overwrite

    def _warmed_up(self) -> bool:
        return bool((self._loss_counts == self.history_per_term).all())

    def weights(self):
        if not self._warmed_up():
            return np.ones([self.T], dtype=np.float64)
        w = np.sqrt(np.mean(self._loss_history ** 2, axis=-1))
        w /= np.sum(w)
        w *= 1 - self.uniform_prob
        w += self.uniform_prob / len(w)
        return w

    def update_with_all_losses(self, ts, losses):
        H = self.history_per_term
        for t, loss in zip(ts, losses):
            t = int(t)
            self._loss_history[t, self._head[t]] = loss
            self._head[t] = (self._head[t] + 1) % H
            if self._loss_counts[t] < H:
                self._loss_counts[t] += 1

    def history(self, t: int) -> np.ndarray:
        """The stored losses of timestep ``t``, oldest first (the layout of the reference's shifted array)."""
        n, h = int(self._loss_counts[t]), int(self._head[t])
        row = self._loss_history[t]
        return row[:n].copy() if n < self.history_per_term else np.concatenate([row[h:], row[:h]])


def create_named_schedule_sampler(name: str, T, *args):
    """core/resample.py:8-22: ``"uniform"`` (T = number of timesteps), ``"snr"`` (T = the per-timestep weight array),
    ``"loss-second-moment"``."""
    if name == "uniform":
        return UniformSampler(T)
    if name == "snr":
        return SNRSampler(T, *args)
    if name == "loss-second-moment":
        return LossSecondMomentResampler(T)
    raise NotImplementedError(f"unknown schedule sampler: {name}")
