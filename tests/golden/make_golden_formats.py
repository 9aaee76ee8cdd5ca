#!/usr/bin/env python3
"""Bundle the reference's own block-compressed test resources (BGZF files and tabix indexes) into one JSON file.

The BGZF / tabix readers of the reference sit on htsjdk (sam-rtg jar) and the JDK's Inflater; neither runs in this image.
Parity of the device decoder is pinned on (a) zlib itself -- Python's `zlib` module is the same library the JDK wraps -- and
(b) the files the reference's own tests read: test/com/rtg/sam/resources/*.gz{,.tbi}, test/com/rtg/tabix/resources/*.gz{,.tbi},
with the known answers of test/com/rtg/tabix/TabixIndexReaderTest.java:50-95 restated in tests/test_formats.py.

/root/reference does not exist on the GPU box, so the files travel as tests/golden/formats_fixtures.json (base64).  Run here with:

    python tests/golden/make_golden_formats.py

Only DATA (test inputs) is bundled -- no reference source code.
"""
import base64
import glob
import json
import os

ROOTS = ["/root/reference/test/com/rtg/sam/resources", "/root/reference/test/com/rtg/tabix/resources"]
MAX_GZ = 120_000   # bytes; the one large file kept (readerWindow1.sam.gz) goes with the TabixIndexReaderTest.testOverlap answer


def main():
    out = {}
    for root in ROOTS:
        for path in sorted(glob.glob(os.path.join(root, "*.gz")) + glob.glob(os.path.join(root, "*.tbi"))):
            name = os.path.basename(path)
            data = open(path, "rb").read()
            if name.endswith(".gz") and len(data) > MAX_GZ:
                continue
            if name.endswith(".gz") and len(data) > 20_000 and name != "readerWindow1.sam.gz":
                continue
            out[name] = base64.b64encode(data).decode()
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "formats_fixtures.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("%d files, %d bytes of JSON" % (len(out), os.path.getsize(dst)))


if __name__ == "__main__":
    main()
