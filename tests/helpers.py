"""Shared test helpers: seeded inputs, bit-slicing, arithmetic ground truth."""
import numpy as np


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def random_sliced(n_rows: int, stride: int, seed: int) -> np.ndarray:
    """Same generator as the device's clg_fill_random: word (row, w) = hi32(splitmix64(seed ^ (row << 40) ^ w))."""
    row = np.arange(n_rows, dtype=np.uint64)[:, None]
    w = np.arange(stride, dtype=np.uint64)[None, :]
    return (splitmix64(np.uint64(seed) ^ (row << np.uint64(40)) ^ w) >> np.uint64(32)).astype(np.uint32)


def stride_for(n_vectors: int) -> int:
    return ((n_vectors + 31) // 32 + 3) // 4 * 4


def exhaustive_sliced(n_rows: int) -> np.ndarray:
    """Vector v (0 <= v < 2^n_rows): immediate i = bit i of v."""
    n = 1 << n_rows
    stride = stride_for(n)
    v = np.arange(stride * 32, dtype=np.uint64)
    out = np.zeros((n_rows, stride), dtype=np.uint32)
    for i in range(n_rows):
        bits = ((v >> np.uint64(i)) & np.uint64(1)).astype(np.uint32).reshape(stride, 32)
        out[i] = (bits << np.arange(32, dtype=np.uint32)).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    return out


def unslice_int(rows: np.ndarray, n_vectors: int) -> np.ndarray:
    """rows [k][stride] -> per-vector integer with bit i = row i (python ints via object dtype when k > 63)."""
    k = rows.shape[0]
    bits = ((rows[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(k, -1)[:, :n_vectors]
    if k <= 63:
        return (bits.astype(np.uint64) << np.arange(k, dtype=np.uint64)[:, None]).sum(axis=0, dtype=np.uint64)
    acc = np.zeros(n_vectors, dtype=object)
    for i in range(k):
        acc += bits[i].astype(object) << i
    return acc


def mask_tail(arr: np.ndarray, n_vectors: int) -> np.ndarray:
    """zero lanes >= n_vectors so buffers compare word for word"""
    arr = arr.copy()
    words = (n_vectors + 31) // 32
    arr[..., words:] = 0
    if n_vectors & 31:
        arr[..., words - 1] &= np.uint32((1 << (n_vectors & 31)) - 1)
    return arr


def sampled_word_columns(n_words: int, k: int = 4, n_ctas: int = 148, extra: int = 24, seed: int = 7) -> np.ndarray:
    """Word columns to check against the oracle at full size: the first and the last column group (k words each), one
    group for EVERY residue of the group index modulo the persistent grid (a CTA of a persistent kernel walks the groups
    g = cta, cta + n_ctas, ...: an indexing error tied to the CTA, the pass number or the tail shows up), and `extra`
    random ones."""
    groups = (n_words + k - 1) // k
    rng = np.random.default_rng(seed)
    pick = {0, groups - 1}
    for r in range(min(n_ctas, groups)):
        pick.add(r + n_ctas * int(rng.integers(0, max(1, (groups - r + n_ctas - 1) // n_ctas))))
    pick |= {int(x) for x in rng.integers(0, groups, size=extra)}
    cols = np.concatenate([np.arange(k * g, min(n_words, k * g + k)) for g in sorted(pick)])
    return cols
