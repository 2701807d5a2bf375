#!/usr/bin/env python3
"""Build the golden fixtures in tests/golden/ from the reference tree.

Run in the build container only (needs /root/reference, PIL); the fixtures it
writes are committed so the tests never read /root/reference at run time.

  aaa_orig_luma512.npz   test-data/aaa-orig.jpg -> PIL .convert('L')
                         .resize((512,512), BILINEAR), exactly the pre-processing
                         of integration/hash.py:65-69 (u8, 512x512).
                         Also the 7 image-space dihedral transforms of that
                         luma plane are derivable by numpy flips, so they are
                         not stored.
  kat.json               the reference's own known answers:
                         - aaa-orig.jpg.expect (3 whitelisted hashes, CI
                           build.yml:96-128)
                         - README doctest hash (README.md:432-436) and PDQF
                           table (README.md:397-412), README-reported
                           aaa-orig.jpg hash/quality/threshold (README.md:164-168)
"""
import json
import os
import re

import numpy as np
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    img = Image.open(os.path.join(REF, "test-data", "aaa-orig.jpg")).convert("L")
    img = img.resize((512, 512), Image.Resampling.BILINEAR)
    luma = np.array(img, dtype=np.uint8)
    assert luma.shape == (512, 512)
    np.savez_compressed(os.path.join(HERE, "aaa_orig_luma512.npz"), luma=luma)

    expect = open(os.path.join(REF, "test-data", "aaa-orig.jpg.expect")).read().split()
    readme = open(os.path.join(REF, "README.md")).read()

    m = re.search(r"let expected = \[(.*?)\];", readme, re.S)
    kat1_hash = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", m.group(1)))
    assert len(kat1_hash) == 32

    m = re.search(r"You should see:\n(.*?)\*/", readme, re.S)
    rows = [r for r in m.group(1).splitlines() if r.strip()]
    table = [[float(x) for x in r.split()] for r in rows]
    assert len(table) == 16 and all(len(r) == 16 for r in table)

    m = re.search(r"Quality: ([0-9.]+)\s+Threshold: ([0-9.]+)\s+Hash: ([0-9a-f]{64})", readme)
    readme_aaa = {"quality": float(m.group(1)), "threshold_avx2": float(m.group(2)),
                  "hash": m.group(3)} if m else None

    kat = {
        "aaa_orig_expect": expect,
        "readme_aaa": readme_aaa,
        "readme_doctest_hash": kat1_hash.hex(),
        "readme_doctest_pdqf_1dp": table,
        "sources": {
            "aaa_orig_expect": "test-data/aaa-orig.jpg.expect:1-3",
            "readme_doctest_input": "README.md:313-322",
            "readme_doctest_hash": "README.md:432-436",
            "readme_doctest_pdqf_1dp": "README.md:397-412",
            "readme_aaa": "README.md:164-168",
        },
    }
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("ok", luma.mean(), kat1_hash.hex(), readme_aaa)


if __name__ == "__main__":
    main()
