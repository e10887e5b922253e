"""Extracts the reference's own golden data for this path into tests/golden/ref_vectors.json.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Sources (paths relative to /root/reference):
    src/constant.rs:1-47,113-114   S_BOX (with its entry-255 deviation), MUL_BY_2, MUL_BY_3,
                                   KEY_SCHEDULE_ROWS, AES_ROWS
    src/utils.rs:28                ROUND_CONSTANT
    src/key_schedule.rs:337-345    EXPANDED (zero-key expansion, 44 words)
    src/utils.rs:40-47             xor_bytes(5, 12) == 9
    src/table.rs:10-16             Tag enum values
Added known-answer vectors that are NOT in the reference (FIPS-197 App. B, App. C.1 and the
zero/zero block) are marked "fips197"; they avoid S-box input 0xFF, where the reference's table
deviates from the standard, so the reference computes true AES on them.
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.json")


def rust_array(text, name):
    m = re.search(r"pub const %s: \[u\d+; \d+\] = \[(.*?)\];" % name, text, re.S)
    return [int(x) for x in re.findall(r"\d+", m.group(1))]


def main():
    const = open(os.path.join(REF, "src/constant.rs")).read()
    utils = open(os.path.join(REF, "src/utils.rs")).read()
    ks = open(os.path.join(REF, "src/key_schedule.rs")).read()
    table = open(os.path.join(REF, "src/table.rs")).read()
    out = {
        "source": "tkmct/halo2-aes (see make_golden.py for file:line)",
        "S_BOX": rust_array(const, "S_BOX"),
        "MUL_BY_2": rust_array(const, "MUL_BY_2"),
        "MUL_BY_3": rust_array(const, "MUL_BY_3"),
        "KEY_SCHEDULE_ROWS": int(re.search(r"KEY_SCHEDULE_ROWS: u64 = (\d+)", const).group(1)),
        "AES_ROWS": int(re.search(r"AES_ROWS: u64 = (\d+)", const).group(1)),
        "ROUND_CONSTANT": [int(x) for x in re.findall(
            r"\d+", re.search(r"ROUND_CONSTANT: \[u64; 10\] = \[(.*?)\]", utils, re.S).group(1))],
        "EXPANDED": re.findall(r'"([0-9a-f]{8})"',
                               re.search(r"const EXPANDED: \[&str; 44\] = \[(.*?)\];", ks, re.S).group(1)),
        "xor_bytes": {"x": 5, "y": 12, "z": 9},
        "Tag": {m.group(1): int(m.group(2)) for m in re.finditer(r"^    (\w+) = (\d+),$", table, re.M)},
        "fips197": [
            {"name": "appendix_B", "key": "2b7e151628aed2a6abf7158809cf4f3c",
             "pt": "3243f6a8885a308d313198a2e0370734", "ct": "3925841d02dc09fbdc118597196a0b32"},
            {"name": "appendix_C1", "key": "000102030405060708090a0b0c0d0e0f",
             "pt": "00112233445566778899aabbccddeeff", "ct": "69c4e0d86a7b0430d8cdb78070b4c55a"},
            {"name": "zero", "key": "00" * 16, "pt": "00" * 16,
             "ct": "66e94bd4ef8a2c3b884cfa59ca342b2e"},
        ],
    }
    assert len(out["S_BOX"]) == 256 and len(out["MUL_BY_2"]) == 256 and len(out["MUL_BY_3"]) == 256
    assert len(out["EXPANDED"]) == 44 and len(out["ROUND_CONSTANT"]) == 10
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
