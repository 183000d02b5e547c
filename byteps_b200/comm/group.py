"""Tiny control-plane abstraction used for bootstrap exchanges (handles,
tokens, plans).  Steady-state synchronisation never goes through here: it is
done by flags in peer-visible device memory."""
import threading


class TorchGroup:
    """Control plane on top of torch.distributed (gloo or nccl default group)."""

    def __init__(self, pg=None):
        import torch.distributed as dist

        self.dist = dist
        self.pg = pg
        self.rank = dist.get_rank(pg)
        self.world = dist.get_world_size(pg)

    def all_gather_object(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.pg)
        return out

    def broadcast_object(self, obj, root=0):
        lst = [obj]
        self.dist.broadcast_object_list(lst, src=root, group=self.pg)
        return lst[0]

    def barrier(self):
        self.dist.barrier(group=self.pg)


class SoloGroup:
    rank = 0
    world = 1

    def all_gather_object(self, obj):
        return [obj]

    def broadcast_object(self, obj, root=0):
        return obj

    def barrier(self):
        pass


class ThreadGroup:
    """In-process group for tests: `world` threads rendezvous on a barrier."""

    class _Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, shared, rank):
        self.shared = shared
        self.rank = rank
        self.world = shared.world

    @staticmethod
    def make(world):
        sh = ThreadGroup._Shared(world)
        return [ThreadGroup(sh, r) for r in range(world)]

    def all_gather_object(self, obj):
        self.shared.slots[self.rank] = obj
        self.shared.barrier.wait()
        out = list(self.shared.slots)
        self.shared.barrier.wait()
        return out

    def broadcast_object(self, obj, root=0):
        return self.all_gather_object(obj)[root]

    def barrier(self):
        self.shared.barrier.wait()
