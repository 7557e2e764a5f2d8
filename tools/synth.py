"""Synthetic .ctx graphs and query batches of the shapes BASELINE.json names.

Tooling, not product: used by tests/ and bench.py to make inputs.  Everything is a pure function of
(seed, index) through a SplitMix64-style counter hash, written with torch int64 ops so the SAME code
produces identical bytes on the CPU (tests, oracle) and on a B200 (bench, large shapes).

Layout produced is the on-disk v6 layout of R/docs/ctx_spec.md (header via the oracle-independent
writer below, records = s LE words (word 0 most significant) + c LE uint32 coverages + c edge bytes).
"""
from __future__ import annotations

import struct

import torch

_M64 = (1 << 64) - 1


def _s64(x: int) -> int:
    """Python int -> the signed 64-bit value with the same bit pattern."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x: torch.Tensor, n: int) -> torch.Tensor:
    """Logical shift right on int64 tensors."""
    if n == 0:
        return x
    return (x >> n) & _s64((1 << (64 - n)) - 1)


def mix64(x: torch.Tensor) -> torch.Tensor:
    """SplitMix64 finaliser (wrapping int64 arithmetic)."""
    x = x + _s64(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _s64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _s64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def hash_idx(seed: int, stream: int, idx: torch.Tensor) -> torch.Tensor:
    return mix64(mix64(idx + _s64(seed * 0x632BE59BD9B4E019 + stream * 0xD1342543DE82EF95)))


def umod(x: torch.Tensor, m: int) -> torch.Tensor:
    """(unsigned 63-bit of x) mod m, as int64 >= 0."""
    return (x & _s64((1 << 63) - 1)) % m


# ------------------------------------------------------------------ packed k-mer arithmetic (word lists)

def _rev2_word(x: torch.Tensor) -> torch.Tensor:
    """Reverse the order of the 32 two-bit groups in each int64."""
    x = (_lsr(x, 2) & _s64(0x3333333333333333)) | ((x & _s64(0x3333333333333333)) << 2)
    x = (_lsr(x, 4) & _s64(0x0F0F0F0F0F0F0F0F)) | ((x & _s64(0x0F0F0F0F0F0F0F0F)) << 4)
    x = (_lsr(x, 8) & _s64(0x00FF00FF00FF00FF)) | ((x & _s64(0x00FF00FF00FF00FF)) << 8)
    x = (_lsr(x, 16) & _s64(0x0000FFFF0000FFFF)) | ((x & _s64(0x0000FFFF0000FFFF)) << 16)
    return _lsr(x, 32) | (x << 32)


def mask_words(words: list[torch.Tensor], k: int) -> list[torch.Tensor]:
    """Zero the unused top bits (bases are right-aligned; top 64s-2k bits must be 0)."""
    s = len(words)
    top_bits = 2 * k - 64 * (s - 1)
    out = list(words)
    if top_bits < 64:
        out[0] = out[0] & _s64((1 << top_bits) - 1)
    return out


def revcomp_words(words: list[torch.Tensor], k: int) -> list[torch.Tensor]:
    """Reverse complement of right-aligned 2-bit k-mers held as s int64 tensors (word 0 most significant)."""
    s = len(words)
    rev = [_rev2_word(~w) for w in reversed(words)]        # sequence now LEFT-aligned in 64s bits
    sh = 64 * s - 2 * k
    if sh == 0:
        return rev
    out = []
    for i in range(s):                                     # logical right shift of the multiword value by sh (<64)
        lo = _lsr(rev[i], sh)
        hi = (rev[i - 1] << (64 - sh)) if i > 0 else torch.zeros_like(lo)
        out.append(lo | hi)
    return out


def _ukey(w: torch.Tensor) -> torch.Tensor:
    """Order-preserving map unsigned-64 -> signed-64."""
    return w ^ _s64(1 << 63)


def less_words(a: list[torch.Tensor], b: list[torch.Tensor]) -> torch.Tensor:
    """Unsigned multiword a < b."""
    lt = torch.zeros_like(a[0], dtype=torch.bool)
    eq = torch.ones_like(a[0], dtype=torch.bool)
    for x, y in zip(a, b):
        lt |= eq & (_ukey(x) < _ukey(y))
        eq &= x == y
    return lt


def canonical_words(words: list[torch.Tensor], k: int):
    rc = revcomp_words(words, k)
    flip = less_words(rc, words)
    return [torch.where(flip, r, w) for w, r in zip(words, rc)], flip


def sort_unique_words(words: list[torch.Tensor]) -> list[torch.Tensor]:
    """Ascending unsigned multiword sort with duplicates removed."""
    perm = None
    for w in reversed(words):                              # LSD: least significant word first, stable
        key = _ukey(w if perm is None else w[perm])
        _, p = torch.sort(key, stable=True)
        perm = p if perm is None else perm[p]
    srt = [w[perm] for w in words]
    if len(srt[0]) > 1:
        same = torch.ones(len(srt[0]) - 1, dtype=torch.bool, device=srt[0].device)
        for w in srt:
            same &= w[1:] == w[:-1]
        keep = torch.cat([torch.ones(1, dtype=torch.bool, device=same.device), ~same])
        srt = [w[keep] for w in srt]
    return srt


def random_words(seed: int, stream: int, n: int, k: int, device, offset: int = 0) -> list[torch.Tensor]:
    s = (k + 31) // 32
    idx = torch.arange(offset, offset + n, dtype=torch.int64, device=device)
    return mask_words([hash_idx(seed, stream * 16 + w, idx) for w in range(s)], k)


def random_canonical_keys(seed: int, n: int, k: int, device) -> list[torch.Tensor]:
    """Exactly n distinct canonical k-mers, ascending, as s int64 word tensors (word 0 most significant)."""
    space = 4 ** k
    if n > space // 3:
        raise ValueError("n too large for k")
    have: list[torch.Tensor] | None = None
    drawn = 0
    for _ in range(64):
        need = n - (0 if have is None else len(have[0]))
        m = int(need * 1.05) + 64
        cand, _ = canonical_words(random_words(seed, 1, m, k, device, offset=drawn), k)
        drawn += m
        have = sort_unique_words(cand if have is None else [torch.cat([a, b]) for a, b in zip(have, cand)])
        if len(have[0]) >= n:
            break
    u = len(have[0])
    if u < n:
        raise RuntimeError("could not draw enough distinct k-mers")
    if u > n:                                              # drop u-n evenly spread elements, order kept
        pick = (torch.arange(n, dtype=torch.int64, device=device) * u) // n
        have = [w[pick] for w in have]
    return have


# ------------------------------------------------------------------ coverage / edges

ADV_PERIOD_DEFAULT = 1_000_003


def coverage_and_edges(seed: int, n: int, c: int, device, novel_permille: int = 5, adv_period: int = ADV_PERIOD_DEFAULT,
                       offset: int = 0):
    """Per-colour uint32 coverage (as int64 holding 0..2^32-1) and edge bytes.
    Colour 0 is the child; colours 1..c-1 are parents / references.  Classes (permille of records):
    shared by all 800-ish, child + one parent 120, parents only 70, child only = NOVEL `novel_permille`,
    child + last colour only 5.  Every adv_period-th record is adversarial:
    variant 0/1 child coverage 0x80000000 / 0xFFFFFFFF with all parents 0 (NOT novel: signed compare),
    variant 2 child 5 and one parent 0x80000000 (blocks novelty: != 0)."""
    idx = torch.arange(offset, offset + n, dtype=torch.int64, device=device)
    r = umod(hash_idx(seed, 2, idx), 1000)
    t_child_last = 1000 - 5
    t_novel = t_child_last - novel_permille
    t_parents = t_novel - 70
    t_child_one = t_parents - 120
    cls_shared = r < t_child_one
    cls_child_one = (r >= t_child_one) & (r < t_parents)
    cls_parents = (r >= t_parents) & (r < t_novel)
    cls_novel = (r >= t_novel) & (r < t_child_last)
    cls_child_last = r >= t_child_last
    which_parent = 1 + umod(hash_idx(seed, 3, idx), max(c - 1, 1))
    cov = torch.zeros((n, c), dtype=torch.int64, device=device)
    edges = torch.zeros((n, c), dtype=torch.int64, device=device)
    for col in range(c):
        if col == 0:
            present = cls_shared | cls_child_one | cls_novel | cls_child_last
        else:
            present = cls_shared | cls_parents | (cls_child_one & (which_parent == col))
            if col == c - 1:
                present = present | cls_child_last
        h = hash_idx(seed, 100 + col, idx)
        cov[:, col] = torch.where(present, 1 + umod(h, 60), torch.zeros_like(h))
        edges[:, col] = torch.where(present, _lsr(h, 40) & 0xFF, torch.zeros_like(h))
    if adv_period > 0:
        adv = (idx % adv_period) == (adv_period // 2)
        variant = (idx // adv_period) % 3
        big = torch.where(variant == 0, torch.full_like(idx, 0x80000000), torch.full_like(idx, 0xFFFFFFFF))
        a01 = adv & (variant < 2)
        a2 = adv & (variant == 2)
        for col in range(c):
            if col == 0:
                cov[:, 0] = torch.where(a01, big, torch.where(a2, torch.full_like(idx, 5), cov[:, 0]))
            else:
                blk = a2 & (which_parent == col) if c > 1 else a2
                cov[:, col] = torch.where(a01, torch.zeros_like(idx), torch.where(blk, torch.full_like(idx, 0x80000000),
                                          torch.where(a2, torch.zeros_like(idx), cov[:, col])))
    return cov, edges


# ------------------------------------------------------------------ record assembly

def assemble_records(words: list[torch.Tensor], cov: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    """[n, 8s+5c] uint8 in on-disk layout."""
    n, c = cov.shape
    s = len(words)
    S = 8 * s + 5 * c
    rec = torch.empty((n, S), dtype=torch.uint8, device=cov.device)
    for w in range(s):
        rec[:, 8 * w:8 * w + 8] = words[w].contiguous().view(torch.uint8).view(n, 8)      # little-endian bytes
    cov32 = cov.to(torch.int64).contiguous().view(torch.uint8).view(n, c, 8)[:, :, :4]
    rec[:, 8 * s:8 * s + 4 * c] = cov32.reshape(n, 4 * c)
    rec[:, 8 * s + 4 * c:] = edges.to(torch.uint8)
    return rec


def default_names(c: int) -> list[str]:
    if c == 4:
        return ["child", "mom", "dad", "ref"]
    return ["child"] + ["p%02d" % i for i in range(c - 1)]


def header_bytes(k: int, c: int, names: list[str] | None = None, mean_read_length: int = 100, graph_name: str = "undefined") -> bytes:
    """v6 header (R/docs/ctx_spec.md tables 1-3); 'undefined' (9 bytes) as McCortex / the fixture write,
    which keeps the data offset off any 16-byte boundary."""
    s = (k + 31) // 32
    names = names or default_names(c)
    out = [b"CORTEX", struct.pack("<4I", 6, k, s, c)]
    out.append(struct.pack("<%dI" % c, *([mean_read_length] * c)))
    out.append(struct.pack("<%dQ" % c, *([0] * c)))
    for nm in names:
        out.append(struct.pack("<I", len(nm)) + nm.encode())
    out.append(bytes([0, 0xD8, 0xA3, 0x70, 0x3D, 0x0A, 0xD7, 0xA3, 0xF8, 0x3F, 0, 0, 0, 0, 0, 0]) * c)
    g = graph_name.encode()
    for _ in range(c):
        out.append(struct.pack("<4B3I", 0, 0, 0, 0, 0, 0, len(g)) + g)
    out.append(b"CORTEX")
    return b"".join(out)


def make_graph_body(seed: int, n: int, k: int, c: int, device="cpu", novel_permille: int = 5,
                    adv_period: int = ADV_PERIOD_DEFAULT, chunk: int = 1 << 24):
    """Returns (body uint8 [n, S] on device, key words list).  Coverage is generated in chunks to bound memory."""
    words = random_canonical_keys(seed, n, k, device)
    s = len(words)
    S = 8 * s + 5 * c
    body = torch.empty((n, S), dtype=torch.uint8, device=device)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        cov, edges = coverage_and_edges(seed, hi - lo, c, device, novel_permille, adv_period, offset=lo)
        body[lo:hi] = assemble_records([w[lo:hi] for w in words], cov, edges)
    return body, words


def make_ctx_file(seed: int, n: int, k: int, c: int, novel_permille: int = 5, adv_period: int = ADV_PERIOD_DEFAULT,
                  names: list[str] | None = None, trailing: bytes = b"") -> bytes:
    """A complete .ctx image (CPU).  `trailing` appends a partial record to exercise the floor in numRecords."""
    body, _ = make_graph_body(seed, n, k, c, "cpu", novel_permille, adv_period)
    return header_bytes(k, c, names) + body.numpy().tobytes() + trailing


# ------------------------------------------------------------------ queries

_ASCII = torch.tensor(list(b"ACGT"), dtype=torch.uint8)


def words_to_ascii(words: list[torch.Tensor], k: int) -> torch.Tensor:
    n = len(words[0])
    s = len(words)
    out = torch.empty((n, k), dtype=torch.uint8, device=words[0].device)
    lut = _ASCII.to(words[0].device)
    for i in range(k):
        bit = 2 * (k - 1 - i)
        code = _lsr(words[s - 1 - bit // 64], bit % 64) & 3
        out[:, i] = lut[code]
    return out


def make_queries(seed: int, table_words: list[torch.Tensor], k: int, nq: int, hit_fraction_permille: int = 500,
                 corrupt_permille: int = 1, offset: int = 0):
    """Uniform-mode queries: hit w.p. ~1/2 (a table k-mer at a hashed index) else a fresh draw;
    random strand; `corrupt_permille` get one base replaced by 'N' and as many are lower-cased (all must miss).
    Returns (ascii [nq,k] uint8, canonical packed words list (zeros where corrupted), expect_valid bool)."""
    dev = table_words[0].device
    n = len(table_words[0])
    idx = torch.arange(offset, offset + nq, dtype=torch.int64, device=dev)
    pick = umod(hash_idx(seed, 7, idx), max(n, 1))
    is_hit = umod(hash_idx(seed, 8, idx), 1000) < hit_fraction_permille
    fresh, _ = canonical_words(random_words(seed, 9, nq, k, dev, offset=offset), k)
    if n > 0:
        canon = [torch.where(is_hit, t[pick], f) for t, f in zip(table_words, fresh)]
    else:
        canon = fresh
    strand = (hash_idx(seed, 10, idx) & 1) == 1
    rc = revcomp_words(canon, k)
    shown = [torch.where(strand, r, w) for w, r in zip(canon, rc)]
    ascii_q = words_to_ascii(shown, k)
    r = umod(hash_idx(seed, 11, idx), 1000)
    put_n = r < corrupt_permille
    lower = (r >= corrupt_permille) & (r < 2 * corrupt_permille)
    pos = umod(hash_idx(seed, 12, idx), k)
    rows = torch.arange(nq, device=dev)
    cur = ascii_q[rows, pos]
    ascii_q[rows, pos] = torch.where(put_n, torch.full_like(cur, ord("N")), cur)
    ascii_q = torch.where(lower[:, None], ascii_q + 32, ascii_q)
    valid = ~(put_n | lower)
    return ascii_q, canon, valid


def random_genome(seed: int, length: int, device="cpu", n_permille: int = 0) -> torch.Tensor:
    """ASCII genome; optionally sprinkle N's."""
    idx = torch.arange(length, dtype=torch.int64, device=device)
    h = hash_idx(seed, 20, idx)
    seq = _ASCII.to(device)[h & 3]
    if n_permille:
        seq = torch.where(umod(hash_idx(seed, 21, idx), 1000) < n_permille, torch.full_like(seq, ord("N")), seq)
    return seq


# ------------------------------------------------------------------ k-mer-range partitioned graphs (BASELINE configs[3])

def canonical_range_bounds(k: int, world: int) -> list[int]:
    """Key-space boundaries that give every rank the same number of canonical k-mers: a k-mer x is canonical iff
    x <= rc(x), and rc(x) is (for this purpose) independent of the leading bases of x, so the density of canonical
    k-mers at relative position t of the key space is 2(1 - t) and the r-th boundary sits at 1 - sqrt(1 - r/world)."""
    import math
    space = 4 ** k
    return [min(space, int(space * (1.0 - math.sqrt(max(0.0, 1.0 - r / world))))) for r in range(world + 1)]


def range_canonical_keys(seed: int, n: int, k: int, lo: int, hi: int, device, chunk: int = 1 << 26) -> list[torch.Tensor]:
    """Exactly n distinct canonical k-mers inside [lo, hi), ascending, one word (k <= 31).  Rejection sampling: uniform
    draws from the range are kept when they are their own canonical form, which samples the canonical k-mers of the
    range uniformly -- every rank builds its own shard of one global sorted graph without seeing the others."""
    assert k <= 31 and 0 <= lo < hi <= 4 ** k
    space = 4 ** k
    acc = max(0.02, 1.0 - (lo + hi) / (2.0 * space))            # expected acceptance = mean canonical density / 2
    have: torch.Tensor | None = None
    drawn = 0
    for _ in range(64):
        got = 0 if have is None else int(have.numel())
        if got >= n:
            break
        want = int((n - got) / acc * 1.02) + 65536
        parts = [] if have is None else [have]
        for o in range(0, want, chunk):
            m = min(chunk, want - o)
            idx = torch.arange(drawn + o, drawn + o + m, dtype=torch.int64, device=device)
            x = umod(hash_idx(seed, 31, idx), hi - lo) + lo
            rc = revcomp_words([x], k)[0]
            parts.append(x[x <= rc])
            del idx, x, rc
        drawn += want
        allk = torch.cat(parts)
        del parts
        srt, _ = torch.sort(allk)
        del allk
        keep = torch.ones(srt.numel(), dtype=torch.bool, device=device)
        keep[1:] = srt[1:] != srt[:-1]
        have = srt[keep]
        del srt, keep
    u = int(have.numel())
    if u < n:
        raise RuntimeError("could not draw enough distinct k-mers in the range")
    if u > n:                                                    # drop u-n evenly spread elements, order kept
        pick = (torch.arange(n, dtype=torch.int64, device=device) * u) // n
        have = have[pick]
    return [have]


def owner_of_keys(words: list[torch.Tensor], splitters: torch.Tensor | None) -> torch.Tensor:
    """Number of splitters <= key (unsigned, lexicographic over the words): the shard that owns the key.  An independent
    torch formulation of the owner rule of the routed lookup, used to check it."""
    n = words[0].numel()
    owner = torch.zeros(n, dtype=torch.int64, device=words[0].device)
    if splitters is None:
        return owner
    for j in range(splitters.shape[0]):
        sp = [splitters[j, w] for w in range(len(words))]
        lt = torch.zeros(n, dtype=torch.bool, device=words[0].device)
        eq = torch.ones(n, dtype=torch.bool, device=words[0].device)
        for w, s in zip(words, sp):                               # key < splitter ?
            lt |= eq & (_ukey(w) < _ukey(s))
            eq &= w == s
        owner += (~lt).to(torch.int64)
    return owner
