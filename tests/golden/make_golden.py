#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors for the k-mer hot path into tests/golden/.

Run HERE (build container, /root/reference present).  The outputs are committed so that neither the
CPU tests nor the GPU tests need /root/reference at run time.

Sources (T/ = public/java/tests/uk/ac/ox/well/cortexjdk/):
  * testdata/two_short_contigs.ctx, .fa            -> copied verbatim (binary test DATA, not source)
  * T/utils/kmer/CortexGraphTest.java:71-136       -> fixture_records.json (66 rows: kmer, cov[2], edges[2])
  * T/utils/kmer/CortexGraphTest.java:322-331      -> the non-existent (N-containing) query
  * T/utils/sequence/SequenceUtilsTest.java:19-57  -> complement table, reverse-complement vectors
  * T/utils/kmer/CanonicalKmerTest.java:8-14       -> hash-colliding pair
  * T/utils/traversal/TraversalEngineTest.java:48-95 -> TempGraphAssembler record strings (k=3, k=5)
"""
import json
import os
import re
import shutil

REF = "/root/reference"
T = os.path.join(REF, "public/java/tests/uk/ac/ox/well/cortexjdk")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    for name in ("two_short_contigs.ctx", "two_short_contigs.fa"):
        shutil.copyfile(os.path.join(REF, "testdata", name), os.path.join(HERE, name))

    src = open(os.path.join(T, "utils/kmer/CortexGraphTest.java")).read()
    row = re.compile(
        r'recs\.add\(new SimpleCortexRecord\("([ACGT]+)",\s*new int\[\]\s*\{\s*(\d+),\s*(\d+)\s*\},'
        r'\s*new String\[\]\s*\{"([^"]{8})",\s*"([^"]{8})"\}\)\);')
    records = [{"kmer": m.group(1), "coverage": [int(m.group(2)), int(m.group(3))],
                "edges": [m.group(4), m.group(5)]} for m in row.finditer(src)]
    assert len(records) == 66, len(records)
    missing = re.search(r'String nonExistentRecord = "([A-Z]+)";', src).group(1)

    seq = open(os.path.join(T, "utils/sequence/SequenceUtilsTest.java")).read()
    trials = re.search(r"byte\[\] trials = new byte\[\] \{([^}]*)\}", seq).group(1)
    exp = re.search(r"byte\[\] exp\s+= new byte\[\] \{([^}]*)\}", seq).group(1)
    comp_in = re.findall(r"'(.)'", trials)
    comp_out = re.findall(r"'(.)'", exp)
    assert len(comp_in) == len(comp_out) == 10
    rc_pairs = re.findall(r'byte\[\] sequence\s+= "([A-Za-z]+)"\.getBytes\(\);\s*byte\[\] expectedRC = "([A-Za-z]+)"', seq)
    assert len(rc_pairs) == 3

    ck = open(os.path.join(T, "utils/kmer/CanonicalKmerTest.java")).read()
    collide = re.findall(r'new CanonicalKmer\("([ACGT]+)"\)', ck)
    assert len(collide) == 2

    te = open(os.path.join(T, "utils/traversal/TraversalEngineTest.java")).read()
    blocks = []
    for name in ("testArbitraryGraphConstruction", "testSlightlyLargerArbitraryGraphConstruction"):
        body = te[te.index("public void " + name):]
        body = body[:body.index("@Test")] if "@Test" in body else body
        haps = re.findall(r'haplotypes\.put\("(\w+)", Collections\.singletonList\("([ACGT]+)"\)\);', body)
        k = int(re.search(r"TempGraphAssembler\.buildGraph\(haplotypes, (\d+)\)", body).group(1))
        n = int(re.search(r"cge\.hasNRecords\((\d+)\)", body).group(1))
        recs = re.findall(r'cge\.hasRecord\("([^"]+)"\);', body)
        assert len(recs) == n
        blocks.append({"haplotypes": haps, "k": k, "records": recs})

    out = {
        "_source": "extracted by tests/golden/make_golden.py from the reference's TestNG sources",
        "fixture_records": records,
        "missing_query": missing,
        "complement_in": comp_in, "complement_out": comp_out,
        "reverse_complement": rc_pairs,
        "hash_collision_pair": collide,
        "assembler_kats": blocks,
    }
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(records), "fixture rows,", [len(b["records"]) for b in blocks], "assembler rows")


if __name__ == "__main__":
    main()
