"""Pair-aware sample scheduling for domain-adaptation training.

  BatchSchedulerSampler             mmdet/datasets/samplers/batch_sampler.py:10-61 - same name, constructor and index
                                    stream: every mini-batch is samples_per_gpu/2 indices of dataset 0 (source) followed
                                    by samples_per_gpu/2 of dataset 1 (target), each dataset shuffled on its own, the
                                    smaller one restarted until the larger one is exhausted.  Used by the reference's
                                    non-distributed loader (mmdet/datasets/builder.py:156-168).
  DistributedBatchSchedulerSampler  the N-GPU form the reference lacks (its distributed branch falls back to
                                    DistributedGroupSampler, which does not keep source/target pairs together, Q14):
                                    the same schedule drawn from a seeded generator that is identical on every rank,
                                    whole mini-batches dealt to the ranks (a source/target pair is never split,
                                    dist.shard_pairs), every rank the same number of mini-batches.

`dataset` is a torch.utils.data.ConcatDataset-like object: `.datasets` (list) and `.cumulative_sizes`.
The labels the loader attaches (`gt_da` = 0 for the source dataset, 1 for the target dataset,
mmdet/datasets/da_dataset.py:105-130) follow from the dataset an index falls into: `domain_of`.
"""
import bisect
import math

import torch
from torch.utils.data import Sampler
from torch.utils.data.sampler import RandomSampler


def domain_of(dataset, index):
    """0 for an index of the first (source) dataset, 1 for the second (target), ... (da_dataset.py:118-122)."""
    return bisect.bisect_right(list(dataset.cumulative_sizes), int(index))


class BatchSchedulerSampler(Sampler):
    def __init__(self, dataset, samples_per_gpu=1):
        self.dataset = dataset
        self.batch_size = int(samples_per_gpu / 2)
        if self.batch_size < 1:
            raise ValueError("samples_per_gpu must hold at least one source and one target image (>= 2)")
        self.number_of_datasets = len(dataset.datasets)
        self.largest_dataset_size = max(len(d) for d in dataset.datasets)

    def __len__(self):
        return self.batch_size * math.ceil(self.largest_dataset_size / self.batch_size) * self.number_of_datasets

    def _schedule(self, make_iter):
        """make_iter(dataset_idx) -> fresh iterator over a permutation of that dataset's local indices."""
        iters = [make_iter(i) for i in range(self.number_of_datasets)]
        first = [0] + list(self.dataset.cumulative_sizes[:-1])
        step = self.batch_size * self.number_of_datasets
        epoch_samples = self.largest_dataset_size * self.number_of_datasets
        out = []
        for _ in range(0, epoch_samples, step):
            for i in range(self.number_of_datasets):
                for _ in range(self.batch_size):
                    try:
                        local = next(iters[i])
                    except StopIteration:            # the smaller dataset starts over
                        iters[i] = make_iter(i)
                        local = next(iters[i])
                    out.append(local + first[i])
        return out

    def __iter__(self):
        samplers = [RandomSampler(d) for d in self.dataset.datasets]      # torch's global RNG, like the reference
        return iter(self._schedule(lambda i: iter(samplers[i])))


class DistributedBatchSchedulerSampler(BatchSchedulerSampler):
    def __init__(self, dataset, samples_per_gpu=2, num_replicas=1, rank=0, seed=0):
        super().__init__(dataset, samples_per_gpu)
        if not 0 <= rank < num_replicas:
            raise ValueError("rank out of range")
        self.num_replicas, self.rank, self.seed, self.epoch = int(num_replicas), int(rank), int(seed), 0
        self.step = self.batch_size * self.number_of_datasets
        total = math.ceil(self.largest_dataset_size * self.number_of_datasets / self.step)
        self.batches_per_rank = math.ceil(total / self.num_replicas)

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def __len__(self):
        return self.batches_per_rank * self.step

    def global_schedule(self):
        """The epoch's mini-batches (lists of `step` indices), identical on every rank."""
        g = torch.Generator()
        g.manual_seed(self.seed + self.epoch)
        sizes = [len(d) for d in self.dataset.datasets]
        flat = self._schedule(lambda i: iter(torch.randperm(sizes[i], generator=g).tolist()))
        batches = [flat[k:k + self.step] for k in range(0, len(flat), self.step)]
        while len(batches) < self.batches_per_rank * self.num_replicas:       # pad by wrapping around: equal work per rank
            batches.append(batches[len(batches) % max(1, len(batches))])
        return batches

    def __iter__(self):
        batches = self.global_schedule()
        mine = batches[self.rank::self.num_replicas][:self.batches_per_rank]
        return iter([i for b in mine for i in b])
