"""Writes the small golden fixtures under tests/golden/.

Sources of the values (none of them can come from running the reference here: it is Rust on un-vendored crates,
no cargo in the image):
  * from_secs_f64.json  -- the documented examples of Rust std `Duration::try_from_secs_f64`
                           (core::time, "the conversion uses rounding with tie resolution to even"), the function
                           the reference calls at discrete.rs:490,860,1107,1380 and delays.rs:102,132,163.
  * philox4x32_10.json  -- Random123 known-answer vectors (kat_vectors) for philox4x32-10.
  * delay_modes.json    -- the reference's own two unit tests, quokkasim/src/delays.rs:182-242, as data.
  * kat0_discrete_queue.json -- the hand-derived trace of SURVEY.md A.8 (E1..E8 + INIT) for discrete_queue.rs with
                           tape P1=[4.0, 5.0], P2=[2.5]: written out by hand below, record by record.
Run: python tests/golden/make_golden.py
"""
import json
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))


def f64_from_bits(b):
    return struct.unpack("<d", struct.pack("<Q", b))[0]


FROM_SECS = [
    # (input, secs, nanos) ; input either a float or {"bits": hex}
    (0.0, 0, 0),
    (1e-20, 0, 0),
    (4.2e-7, 0, 420),
    (2.7, 2, 700000000),
    ({"bits": "0x0000000000000001"}, 0, 0),  # subnormal
    (0.999e-9, 0, 1),
    (0.999999999499, 0, 999999999),
    (0.999999999501, 1, 0),
    (42.999999999499, 42, 999999999),
    (42.999999999501, 43, 0),
    ({"bits": "0x3F50000000000000"}, 0, 976562),  # exactly 976562.5e-9 -> ties to even
    ({"bits": "0x3F68000000000000"}, 0, 2929688),  # exactly 2929687.5e-9
    ({"bits": "0x3FF0040000000000"}, 1, 976562),  # exactly 1.0009765625
    ({"bits": "0x3FF00C0000000000"}, 1, 2929688),  # exactly 1.0029296875
]
# Rust errors: negative, NaN, overflow (>= 2^64 s).  3e10 s is Ok(30_000_000_000 s) in Rust but exceeds this
# executor's u64-nanosecond time base (584 years), so it is a fault here too -- documented in DESIGN.md.
FROM_SECS_ERR = [-1.0, -5e-11, float("nan"), float("inf"), 2e19, 3e10]

PHILOX = [
    {"ctr": ["00000000"] * 4, "key": ["00000000"] * 2, "out": ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]},
    {"ctr": ["ffffffff"] * 4, "key": ["ffffffff"] * 2, "out": ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]},
    {"ctr": ["243f6a88", "85a308d3", "13198a2e", "03707344"], "key": ["a4093822", "299f31d0"],
     "out": ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]},
]

# delays.rs:182-242 -- modes are (until_delay Constant, until_fix Constant); steps are
# ("update", dt_secs, from, to) or ("next", expected_secs); from/to are mode names or None
DELAY_MODES = {
    "test_transition_for_single_delay": {
        "modes": [["TestDelay", 13.0, 5.0]],
        "steps": [["update", 4, None, None], ["update", 10, None, "TestDelay"],
                  ["update", 1, "TestDelay", "TestDelay"], ["update", 5, "TestDelay", None]],
    },
    "test_transition_for_multiple_delays_and_time_to_next": {
        "modes": [["Delays1", 13.0, 5.0], ["Delays2", 15.0, 4.0]],
        "steps": [["next", 13], ["update", 13, None, "Delays1"], ["next", 5], ["update", 5, "Delays1", None],
                  ["next", 2], ["update", 2, None, "Delays2"], ["next", 4], ["update", 4, "Delays2", None],
                  ["next", 11], ["update", 11, None, "Delays1"]],
    },
}

# SURVEY.md A.8.  Components: 0 Source, 1 Queue1, 2 Process1, 3 Queue2, 4 Process2, 5 Queue3, 6 Sink.
# record = [time_ns, comp, kind, reason, seq, src_comp, src_seq, payload]; kinds: 1 Start 3 Finish 4 NonStart
# 6 WithdrawRequest 16 Add 17 Remove 18 StateChange; reason 1 = "Upstream is empty"; StateChange reason = category
# (0 Empty 1 Normal 2 Full), payload = occupied<<32 | empty; item handle = index (single source, ordinal 0).
INIT = 0xFFFF
S = 10**9
KAT0 = {
    "tapes": {"Process1": [4.0, 5.0, 3.0, 3.0], "Process2": [2.5, 2.5, 2.5]},
    "n_events": 8,
    "records": [
        # INIT t=0: SRC#0 ProcessStart(Item_0000); P1#0, P2#0, SNK#0 ProcessNonStart("Upstream is empty")
        [0, 0, 1, 0, 0, INIT, 0, 0],
        [0, 2, 4, 1, 0, INIT, 0, 0],
        [0, 4, 4, 1, 0, INIT, 0, 0],
        [0, 6, 4, 1, 0, INIT, 0, 0],
        # E1 3.000000000 SRC keyed: SRC#1 Finish(Item_0000); Q1#0 Add; SRC#2 Start(Item_0001)
        [3 * S, 0, 3, 0, 1, 0, 0, 0],
        [3 * S, 1, 16, 0, 0, 0, 1, 0],
        [3 * S, 0, 1, 0, 2, 0, 1, 1],
        # E2 3.000000001 Q1 emit: Q1#1 StateChange(Normal{1,9}); P1#1 WithdrawRequest; Q1#2 Remove(Item_0000);
        #    P1#2 Start(Item_0000)
        [3 * S + 1, 1, 18, 1, 1, 1, 0, (1 << 32) | 9],
        [3 * S + 1, 2, 6, 0, 1, 1, 1, 0],
        [3 * S + 1, 1, 17, 0, 2, 2, 1, 0],
        [3 * S + 1, 2, 1, 0, 2, 2, 1, 0],
        # E3 3.000000002 Q1 emit: Q1#3 StateChange(Empty{0,10})
        [3 * S + 2, 1, 18, 0, 3, 1, 2, 10],
        # E4 6.000000000 SRC keyed: SRC#3 Finish(Item_0001); Q1#4 Add; SRC#4 Start(Item_0002)
        [6 * S, 0, 3, 0, 3, 0, 2, 1],
        [6 * S, 1, 16, 0, 4, 0, 3, 1],
        [6 * S, 0, 1, 0, 4, 0, 3, 2],
        # E5 6.000000001 Q1 emit: Q1#5 StateChange(Normal{1,9}) (P1 busy: no logs)
        [6 * S + 1, 1, 18, 1, 5, 1, 4, (1 << 32) | 9],
        # E6 7.000000001 P1 keyed: P1#3 Finish(Item_0000); Q2#0 Add; P1#4 WithdrawRequest; Q1#6 Remove(Item_0001);
        #    P1#5 Start(Item_0001)
        [7 * S + 1, 2, 3, 0, 3, 2, 2, 0],
        [7 * S + 1, 3, 16, 0, 0, 2, 3, 0],
        [7 * S + 1, 2, 6, 0, 4, 2, 3, 0],
        [7 * S + 1, 1, 17, 0, 6, 2, 4, 1],
        [7 * S + 1, 2, 1, 0, 5, 2, 4, 1],
        # E7 7.000000002 Q1 emit (rank 1 first): Q1#7 StateChange(Empty)
        [7 * S + 2, 1, 18, 0, 7, 1, 6, 10],
        # E8 7.000000002 Q2 emit: Q2#1 StateChange(Normal{1,9}); P2#1 WithdrawRequest; Q2#2 Remove(Item_0000);
        #    P2#2 Start(Item_0000)
        [7 * S + 2, 3, 18, 1, 1, 3, 0, (1 << 32) | 9],
        [7 * S + 2, 4, 6, 0, 1, 3, 1, 0],
        [7 * S + 2, 3, 17, 0, 2, 4, 1, 0],
        [7 * S + 2, 4, 1, 0, 2, 4, 1, 0],
    ],
}


def main():
    fs = []
    for x, s, n in FROM_SECS:
        if isinstance(x, dict):
            fs.append({"bits": x["bits"], "ns": s * S + n})
        else:
            fs.append({"bits": "0x%016x" % struct.unpack("<Q", struct.pack("<d", x))[0], "value": repr(x),
                       "ns": s * S + n})
    errs = ["0x%016x" % struct.unpack("<Q", struct.pack("<d", x))[0] for x in FROM_SECS_ERR]
    json.dump({"ok": fs, "error_bits": errs}, open(os.path.join(HERE, "from_secs_f64.json"), "w"), indent=1)
    json.dump(PHILOX, open(os.path.join(HERE, "philox4x32_10.json"), "w"), indent=1)
    json.dump(DELAY_MODES, open(os.path.join(HERE, "delay_modes.json"), "w"), indent=1)
    json.dump(KAT0, open(os.path.join(HERE, "kat0_discrete_queue.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
