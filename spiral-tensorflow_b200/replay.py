"""``ReplayBuffer`` -- the reference's replay.py:8-37 as a device-resident uint8 ring of final
canvases feeding the discriminator's "fake" batch.  Same semantics: pushes truncate
float -> uint8; a push that does not fit shifts the buffer left by the batch size and writes the batch
at the end WITHOUT moving ``idx`` (replay.py:24-27), so sampling stays uniform over ``[0, idx)`` --
rows at or above ``idx`` are sampled only once later pushes have shifted them down (a quirk of the
reference when the push size does not divide the buffer size; kept, not fixed).  Deviations: ``sample``
raises on an empty buffer instead of spinning (replay.py:33-34 waits for the actor threads) and draws
``n`` rows (the reference draws ``disc_batch_size`` whatever ``n`` is; every caller passes that value);
a push larger than the buffer keeps its last ``buffer_size`` rows (the reference would fail to
broadcast).  tests/test_cpu_host.py replays random push sequences against a literal numpy restatement."""
import torch


class ReplayBuffer(object):
    def __init__(self, args, observation_shape, device=None):
        self.buffer_size = args.buffer_batch_num * args.disc_batch_size
        self.batch_size = args.disc_batch_size
        self.device = torch.device("cuda") if device is None else torch.device(device)
        self.rng = torch.Generator(device=self.device)
        self.rng.manual_seed(int(args.seed))
        self.idx = 0
        self.data = torch.zeros([self.buffer_size] + list(observation_shape), dtype=torch.uint8, device=self.device)

    def push(self, batches):
        batches = batches.to(self.device).to(torch.uint8)
        n = batches.shape[0]
        if n > self.buffer_size:
            batches, n = batches[-self.buffer_size:], self.buffer_size
        if self.idx + n > self.buffer_size:
            self.data[:-n] = self.data[n:].clone()
            self.data[-n:] = batches
        else:
            self.data[self.idx:self.idx + n] = batches
            self.idx += n

    def sample(self, n=None):
        n = self.batch_size if n is None else n
        if self.idx < 1:
            raise RuntimeError("replay buffer is empty")
        i = torch.randint(0, self.idx, (n,), generator=self.rng, device=self.device)
        return self.data[i].to(torch.float32)
