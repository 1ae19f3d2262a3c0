"""Character tokenizer, batch sampler and train/val split (reference: src/preprocessing.py).

``get_mapper`` is bit-exact with the reference (sorted set of characters).  ``get_batch`` keeps
the reference's host-side sampler (same torch CPU RNG consumption, so seeded runs draw the same
windows); ``DeviceBatcher`` is the B200-side replacement that keeps the corpus resident in HBM.
"""
import torch


def get_mapper(text):
    """encode / decode / vocab_size for the characters of ``text`` (src/preprocessing.py:3-26)."""
    alphabet = sorted(set(text))
    index_of = {ch: i for i, ch in enumerate(alphabet)}

    def encode(s):
        return [index_of[ch] for ch in s]

    def decode(ids):
        return "".join(alphabet[int(i)] for i in ids)

    return encode, decode, len(alphabet)


def get_batch(data, context_length, batch_size, device):
    """Random (x, y) windows, y shifted by one (src/preprocessing.py:28-46).

    One ``torch.randint`` call on the CPU generator like the reference; the windows are
    gathered with a single indexed read instead of 2*B Python slices.
    """
    ix = torch.randint(len(data) - context_length, (batch_size,))
    offs = ix.unsqueeze(1) + torch.arange(context_length + 1).unsqueeze(0)
    win = data[offs.to(data.device)]
    x, y = win[:, :-1].contiguous(), win[:, 1:].contiguous()
    return x.to(device, non_blocking=True), y.to(device, non_blocking=True)


class DeviceBatcher:
    """Corpus resident on the GPU; windows drawn on the device (CUDA-graph friendly)."""

    def __init__(self, data, context_length, batch_size, device, seed=0):
        self.data = data.to(device)
        self.T, self.B = context_length, batch_size
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)
        self.ar = torch.arange(context_length + 1, device=device).unsqueeze(0)

    def next(self):
        ix = torch.randint(len(self.data) - self.T, (self.B, 1), device=self.data.device, generator=self.gen)
        win = self.data[ix + self.ar]
        return win[:, :-1].contiguous(), win[:, 1:].contiguous()


def get_train_val_data(input_path, train_path, val_path):
    """Encode the corpus, split 90/10, save both tensors (src/preprocessing.py:48-86)."""
    torch.manual_seed(42)
    with open(input_path, "r", encoding="utf-8") as f:
        text = f.read()
    encode, decode, vocab_size = get_mapper(text)
    print(f"Vocab size of the text: {vocab_size}")
    data = torch.tensor(encode(text), dtype=torch.long)
    n = len(data)
    train_data, val_data = data[: int(0.9 * n)], data[int(0.9 * n):]
    x, y = get_batch(train_data, 8, 4, torch.device("cpu"))
    print(f"Input (encoded):\n{x[0].tolist()}\nInput (decoded):\n{decode(x[0].tolist())}\n"
          f"Output (encoded):\n{y[0].tolist()}\nOutput (decoded):\n{decode(y[0].tolist())}\n")
    torch.save(train_data, train_path)
    torch.save(val_data, val_path)
