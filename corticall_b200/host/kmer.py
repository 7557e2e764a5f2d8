"""Host-side value types of the k-mer path, mirroring the reference's Java classes one to one.

These are VALUE TYPES (one k-mer at a time, used as dictionary keys and in `toString`), exactly as in the
reference where they stay plain Java objects on the host.  Nothing here is a
substitute for the CUDA path: every batched operation (record decode, novelty scan, canonicalise+pack of
sequences, lookups) goes through libcorticall_cuda -- see cortex.py.

Reference (S/ = public/java/src/uk/ac/ox/well/cortexjdk/):
  S/utils/sequence/SequenceUtils.java:61-86,127-135,206-225,727-736
  S/utils/kmer/CanonicalKmer.java:8-98, CortexByteKmer.java:11-55, CortexBinaryKmer.java:9-52
  S/utils/io/graph/cortex/CortexRecord.java:291-360 (encode/decodeBinaryKmer)
"""
from __future__ import annotations

_COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")     # everything else (N, n, '.', ...) maps to itself


def _java_array_hash(b: bytes) -> int:
    """java.util.Arrays.hashCode(byte[]) with 32-bit wrap (bytes are signed)."""
    h = 1
    for x in b:
        h = (31 * h + (x - 256 if x > 127 else x)) & 0xFFFFFFFF
    return h - (1 << 32) if h >= (1 << 31) else h


class SequenceUtils:
    """The four SequenceUtils members on the hot path."""

    @staticmethod
    def complement(b):                                     # SequenceUtils.java:61-86
        if isinstance(b, int):
            return _COMP[b]
        return bytes(b).translate(_COMP)

    @staticmethod
    def reverseComplement(seq) -> bytes:                   # :127-135
        if isinstance(seq, str):
            return bytes(seq.encode("latin-1")).translate(_COMP)[::-1].decode("latin-1")
        return bytes(seq).translate(_COMP)[::-1]

    @staticmethod
    def alphanumericallyLowestOrientation(seq):            # :206-225 (signed-byte compare, tie -> forward)
        is_str = isinstance(seq, str)
        b = seq.encode("latin-1") if is_str else bytes(seq)
        rc = b.translate(_COMP)[::-1]
        out = b
        for f, r in zip(b, rc):
            sf, sr = (f - 256 if f > 127 else f), (r - 256 if r > 127 else r)
            if sf < sr:
                break
            if sf > sr:
                out = rc
                break
        return out.decode("latin-1") if is_str else out

    @staticmethod
    def kmerizeSequence(seq: str, kmer_size: int):         # :727-736
        return [CanonicalKmer(seq[i:i + kmer_size]) for i in range(0, len(seq) - kmer_size + 1)]


class CortexByteKmer:
    """ASCII k-mer with the signed-byte lexicographic compareTo the reference's binary search uses."""

    __slots__ = ("kmer",)

    def __init__(self, kmer):
        self.kmer = kmer.encode("latin-1") if isinstance(kmer, str) else bytes(kmer)

    def length(self) -> int:
        return len(self.kmer)

    def getKmer(self) -> bytes:
        return self.kmer

    def setKmer(self, kmer) -> None:
        self.kmer = bytes(kmer)

    def compareTo(self, o: "CortexByteKmer") -> int:       # CortexByteKmer.java:41-49: over THIS length
        other = o.kmer
        for i in range(len(self.kmer)):
            a, b = self.kmer[i], other[i]                  # IndexError == Java's ArrayIndexOutOfBounds
            a, b = (a - 256 if a > 127 else a), (b - 256 if b > 127 else b)
            if a < b:
                return -1
            if a > b:
                return 1
        return 0

    def __eq__(self, o):
        return isinstance(o, CortexByteKmer) and self.kmer == o.kmer

    def __hash__(self):
        return _java_array_hash(self.kmer)

    def __lt__(self, o):
        return self.compareTo(o) < 0

    def __str__(self):
        return self.kmer.decode("latin-1")

    toString = __str__
    hashCode = __hash__


class CanonicalKmer:
    """Lexicographically lowest orientation of a k-mer (CanonicalKmer.java:13-37)."""

    __slots__ = ("kmer", "sk", "_flipped")

    def __init__(self, kmer, kmerIsAlphanumericallyLowest: bool = False):
        b = kmer.encode("latin-1") if isinstance(kmer, str) else bytes(kmer)
        self._flipped = False
        if kmerIsAlphanumericallyLowest:
            self.kmer = b
        else:
            self.kmer = SequenceUtils.alphanumericallyLowestOrientation(b)
            # The reference decides "flipped" by comparing Arrays.hashCode of the two arrays (:16,:23,:33),
            # which is wrong on hash collisions (CanonicalKmerTest.java:8-14); mirrored as is.
            self._flipped = _java_array_hash(self.kmer) != _java_array_hash(b)
        self.sk = self.kmer.decode("latin-1")

    def length(self) -> int:
        return len(self.kmer)

    __len__ = length

    def charAt(self, i: int) -> str:
        return chr(self.kmer[i])

    def subSequence(self, start: int, end: int) -> "CanonicalKmer":
        return self.getSubKmer(start, end - start)

    def isFlipped(self) -> bool:
        return self._flipped

    def getKmerAsBytes(self) -> bytes:
        return self.kmer

    def getKmerAsString(self) -> str:
        return self.sk

    def getSubKmer(self, start: int, length: int) -> "CanonicalKmer":
        return CanonicalKmer(self.kmer[start:start + length])

    def __hash__(self):
        return _java_array_hash(self.kmer)

    hashCode = __hash__

    def __eq__(self, o):                                   # :77-87
        if isinstance(o, CanonicalKmer):
            return self.kmer == o.kmer
        if isinstance(o, str):
            return self.kmer == o.encode("latin-1")
        if isinstance(o, (bytes, bytearray)):
            return self.kmer == bytes(o)
        return False

    equals = __eq__

    def compareTo(self, o: "CanonicalKmer") -> int:        # String.compareTo
        a, b = self.sk, o.sk
        return (a > b) - (a < b)

    def __lt__(self, o):
        return self.sk < o.sk

    def __str__(self):
        return self.sk

    toString = __str__


def _swap64(x: int) -> int:
    return int.from_bytes((x & 0xFFFFFFFFFFFFFFFF).to_bytes(8, "little"), "big")


def _to_signed(x: int) -> int:
    return x - (1 << 64) if x >= (1 << 63) else x


_CODE = {65: 0, 97: 0, 67: 1, 99: 1, 71: 2, 103: 2, 84: 3, 116: 3}


def getKmerBits(kmerSize: int) -> int:                     # CortexRecord.java:309-311
    return (kmerSize + 31) // 32


def native_words(kmer: bytes) -> list[int]:
    """k-mer -> s unsigned native words (word 0 most significant, bases right-aligned) = the on-disk words."""
    k = len(kmer)
    v = 0
    for ch in kmer:
        if ch not in _CODE:
            raise RuntimeError("Nucleotide '%s' is not a valid nucleotide" % chr(ch))   # CortexRecord.java:347-360
        v = (v << 2) | _CODE[ch]
    s = getKmerBits(k)
    return [(v >> (64 * (s - 1 - i))) & 0xFFFFFFFFFFFFFFFF for i in range(s)]


def encodeBinaryKmer(kmer: bytes) -> list[int]:
    """CortexRecord.encodeBinaryKmer :313-334 -- Java long[] convention: BYTE-SWAPPED on-disk words, signed."""
    return [_to_signed(_swap64(w)) for w in native_words(bytes(kmer))]


def decodeBinaryKmer(binaryKmer, kmerSize: int, kmerBits: int) -> bytes:
    """CortexRecord.decodeBinaryKmer :291-307 (input in the Java long[] convention)."""
    v = 0
    for w in binaryKmer[:kmerBits]:
        v = (v << 64) | _swap64(int(w))
    out = bytearray(kmerSize)
    for i in range(kmerSize - 1, -1, -1):
        out[i] = b"ACGT"[v & 3]
        v >>= 2
    return bytes(out)


class CortexBinaryKmer:
    """Packed k-mer in the Java long[] convention (CortexBinaryKmer.java:9-52)."""

    __slots__ = ("binaryKmer",)

    def __init__(self, kmer):
        if isinstance(kmer, (bytes, bytearray, str)):      # (byte[]) ctor canonicalises first (:15-17)
            b = kmer.encode("latin-1") if isinstance(kmer, str) else bytes(kmer)
            self.binaryKmer = tuple(encodeBinaryKmer(SequenceUtils.alphanumericallyLowestOrientation(b)))
        else:
            self.binaryKmer = tuple(int(x) for x in kmer)

    def getBinaryKmer(self):
        return list(self.binaryKmer)

    def __eq__(self, o):
        return isinstance(o, CortexBinaryKmer) and self.binaryKmer == o.binaryKmer

    def __hash__(self):
        h = 1
        for x in self.binaryKmer:
            x &= 0xFFFFFFFFFFFFFFFF
            h = (31 * h + ((x ^ (x >> 32)) & 0xFFFFFFFF)) & 0xFFFFFFFF
        return h - (1 << 32) if h >= (1 << 31) else h

    def compareTo(self, o: "CortexBinaryKmer") -> int:     # :41-51: SIGNED compare of the swapped longs (not lexicographic!)
        for a, b in zip(self.binaryKmer, o.binaryKmer):
            if a != b:
                return -1 if a < b else 1
        return 0
