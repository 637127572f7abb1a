#!/usr/bin/env python
"""Generates arks_b200/csrc/json_tables.h: the transition tables of the gateway's JSON automaton.

The device parses one document per lane, 32 lanes in lock step. A switch-based state machine makes every lane pay for
every other lane's branch (ncu, round 1: 9.6 of 32 threads active on SSE traffic), so the automaton is table driven:

    t = T[state][class(byte)]     (one byte)      t < 240: next state        t >= 240: an event the engine handles

To keep events rare the CONTEXT of a value is part of the state: a string / literal / number that is an array element
(context A), a member of an ordinary object (O), of a jsoniter struct level (S), of an object whose keys are matched
exactly (X: the usage object, the top level of an SSE event) or the top-level value itself (T) has its own copy of the
value states, so finishing it leads straight to the right "after value" state without consulting the container stack.
Events remain for: brackets (the stack), keys of S / X objects (hashing, dispatch), the first byte of S / X / T values
(the members the gateway reads), the end of strings / numbers in S / X (captured values) and errors.

Flavors:
  J  json-iterator v1.1.12 ConfigFastest: struct decoding at the top (readObjectStart / readFieldHash: keys must be
     strings, no control-character check in keys), strict Skip() below (ReadObjectCB keys read with ReadString: the
     literal null is accepted as a key, control characters rejected only before a string's first backslash,
     trySkipNumber leniency), a NUL byte after the document ends parsing successfully.
  E  encoding/json checkValid: RFC 8259.
The number states are found by breadth-first search over an executable copy of the trySkipNumber x RFC-number-DFA
product that oracle/ork_json.c restates (divergence D1), so they cannot drift from it.

Usage: python tools/gen_json_tables.py   (rewrites the header; tests/test_tables_fresh.py checks it is up to date)
"""
import os

CLASSES = ["OTHER", "SP", "WSC", "CTRL", "NUL", "QUOTE", "BSLASH", "COMMA", "COLON", "LBRACE", "RBRACE", "LBRACK", "RBRACK",
           "MINUS", "PLUS", "DOT", "ZERO", "DIG19", "e", "E", "n", "t", "f", "u", "l", "r", "a", "s", "b", "SLASH",
           "HEXLO", "HEXUP"]
CI = {n: i for i, n in enumerate(CLASSES)}
NCLS = 32
assert len(CLASSES) == NCLS


def cls_of(b):
    c = chr(b)
    if b == 0: return "NUL"
    if c == " ": return "SP"
    if c in "\t\n\r": return "WSC"
    if b < 0x20: return "CTRL"
    m = {'"': "QUOTE", "\\": "BSLASH", ",": "COMMA", ":": "COLON", "{": "LBRACE", "}": "RBRACE", "[": "LBRACK",
         "]": "RBRACK", "-": "MINUS", "+": "PLUS", ".": "DOT", "0": "ZERO", "e": "e", "E": "E", "n": "n", "t": "t",
         "f": "f", "u": "u", "l": "l", "r": "r", "a": "a", "s": "s", "b": "b", "/": "SLASH"}
    if c in m: return m[c]
    if c in "123456789": return "DIG19"
    if c in "cd": return "HEXLO"
    if c in "ABCDF": return "HEXUP"
    return "OTHER"


HEX = {"ZERO", "DIG19", "a", "b", "HEXLO", "e", "f", "HEXUP", "E"}
DIGIT = {"ZERO", "DIG19"}
WS = {"SP", "WSC"}
CTRLS = {"WSC", "CTRL", "NUL"}
ESC_OK = {"QUOTE", "BSLASH", "SLASH", "b", "f", "n", "r", "t"}

# events (table values >= EV_BASE)
EV_BASE = 240
EVENTS = ["ERR", "PUSHO", "PUSHA", "POP", "KEY_BEGIN", "KEY_END", "VALUE_BEGIN", "STR_DONE", "NUM_DONE", "CHOICE_ELEM",
          "TOP_OBJ"]
EV = {n: EV_BASE + i for i, n in enumerate(EVENTS)}

# ---- executable copy of the number semantics (oracle j_skip_number + rfc_number_end)
FM, FZ, FI, FD, FF, FE, FS, FX, DEAD = range(9)
ACCEPT = {FZ, FI, FF, FX}
TERM = {"COMMA", "RBRACK", "RBRACE", "SP", "WSC"}
NUMBYTE = DIGIT | {"DOT", "e", "E", "PLUS", "MINUS"}


def num_step(state, k):
    tsn, any_, dot, need, nf = state
    if tsn:
        if need:
            if k not in DIGIT: return ("err",)
            need = 0
        elif k in DIGIT:
            pass
        elif k == "DOT":
            if dot: return ("err",)
            dot, need = 1, 1
        elif k in TERM:
            if any_: return ("done_re",)
            tsn = 0
        else:
            tsn = 0
        any_ = 1
    if k not in NUMBYTE:
        return ("done_re",) if nf in ACCEPT else ("err",)
    nx = DEAD
    if k in DIGIT:
        if nf == FM: nx = FZ if k == "ZERO" else FI
        elif nf == FI: nx = FI
        elif nf in (FD, FF): nx = FF
        elif nf in (FE, FS, FX): nx = FX
    elif k == "DOT":
        if nf in (FZ, FI): nx = FD
    elif k in ("e", "E"):
        if nf in (FZ, FI, FF): nx = FE
    else:
        if nf == FE: nx = FS
    if nx == DEAD and not tsn: return ("err",)
    if not tsn: any_, dot, need = 0, 0, 0
    return ("next", (tsn, any_, dot, need, nx))


def num_start(flavor, k):
    tsn = 1 if (flavor == "J" and k != "ZERO") else 0
    return (tsn, 0, 0, 0, FM if k == "MINUS" else FZ if k == "ZERO" else FI)


# string families: (name, reads-like, after-state)   reads-like: "string" = ReadString / encoding/json, "field" = readFieldHash
def build(flavor):
    ctxs = ["A", "O", "S", "X", "T"]
    names = []
    sid = {}

    def S(n):
        if n not in sid:
            sid[n] = len(names)
            names.append(n)
        return n

    # 1) bulk-skippable string states first, in (plain, _E) pairs so that (state & 1) == "a backslash was seen"
    fams = ["V_" + c for c in ctxs] + ["K", "S", "X"]  # value strings per context, generic / struct / exact keys
    for f in fams:
        S("STR" + f)
        S("STR" + f + "_E")
    n_str = len(names)
    fixed_tokens = ["TOP", "ARR_FIRST", "ARRC_FIRST", "OBJ_FIRST", "OBJ_KEY", "OBJX_FIRST", "OBJX_KEY", "STRUCT_FIRST",
                    "STRUCT_KEY", "COLON_O", "COLON_S", "COLON_X", "AFTER_A", "AFTER_O", "AFTER_S", "AFTER_X", "FINISH",
                    "STOP", "ERRSTATE"] + ["VAL_" + c for c in ctxs] + ["VALG_S", "VALG_X", "VALG_T"]
    for t in fixed_tokens:
        S(t)
    for f in fams:
        for e in ("ESC", "U4", "U3", "U2", "U1"):
            S(e + f)
    lit_chain = {"N1": ("u", "N2"), "N2": ("l", "N3"), "N3": ("l", None), "T1": ("r", "T2"), "T2": ("u", "T3"), "T3": ("e", None),
                 "F1": ("a", "F2"), "F2": ("l", "F3"), "F3": ("s", "F4"), "F4": ("e", None)}
    for c in ctxs:
        for l in lit_chain:
            S(l + "_" + c)
    for l in ("NK1", "NK2", "NK3", "NKX1", "NKX2", "NKX3"):
        S(l)

    after = {"A": "AFTER_A", "O": "AFTER_O", "S": "AFTER_S", "X": "AFTER_X", "T": "FINISH"}
    rows = {}
    work = []

    def num(st, c):
        n = "NUM_%s_%d%d%d%d_%d" % ((c,) + st)
        if n not in sid:
            S(n)
            work.append((st, c, n))
        return n

    def token_row(s):
        """transitions of a between-tokens state, as dict class -> target (state name or event name)"""
        r = {}
        for k in CLASSES:
            r[k] = "ERR"
        for k in WS:
            r[k] = s
        return r

    def value_start(c, k, generic_ok=True):
        """first byte of a value in context c (generic handling)"""
        if k == "QUOTE": return "STRV_" + c
        if k == "n": return "N1_" + c
        if k == "t": return "T1_" + c
        if k == "f": return "F1_" + c
        if k in ("MINUS", "ZERO", "DIG19"): return num(num_start(flavor, k), c)
        if k == "LBRACK": return "PUSHA"
        if k == "LBRACE": return "PUSHO"
        return "ERR"

    def after_row(c):
        r = token_row(after[c])
        if c == "A":
            r["COMMA"] = "VAL_A"
            r["RBRACK"] = "POP"
        elif c == "O":
            r["COMMA"] = "OBJ_KEY"
            r["RBRACE"] = "POP"
        elif c == "S":
            r["COMMA"] = "STRUCT_KEY"
            r["RBRACE"] = "POP"
        elif c == "X":
            r["COMMA"] = "OBJX_KEY"
            r["RBRACE"] = "POP"
        else:  # FINISH
            if flavor == "J": r["NUL"] = "STOP"  # frozenConfig.Unmarshal: `if c == 0` also matches a NUL byte
        return r

    # token states
    r = token_row("TOP"); r["LBRACE"] = "TOP_OBJ"; r["n"] = "N1_T"; rows["TOP"] = r  # jsoniter readObjectStart
    for c in ctxs:
        r = token_row("VAL_" + c)
        for k in CLASSES:
            if k in WS: continue
            if c in ("S", "X", "T") and not (c == "T" and flavor == "J"):
                r[k] = "VALUE_BEGIN"  # the engine looks at the pending member first, then re-dispatches in VALG_c
            else:
                r[k] = value_start(c, k)
        rows["VAL_" + c] = r
    for c in ("S", "X", "T"):
        r = token_row("VALG_" + c)
        for k in CLASSES:
            if k in WS: continue
            r[k] = value_start(c, k)
        rows["VALG_" + c] = r
    for nm, first_is_choice in (("ARR_FIRST", False), ("ARRC_FIRST", True)):
        r = token_row(nm)
        for k in CLASSES:
            if k in WS: continue
            r[k] = "CHOICE_ELEM" if first_is_choice else value_start("A", k)
        r["RBRACK"] = "POP"
        rows[nm] = r
    r = token_row("OBJ_FIRST"); r["QUOTE"] = "STRK"; r["RBRACE"] = "POP"; rows["OBJ_FIRST"] = r
    r = token_row("OBJ_KEY"); r["QUOTE"] = "STRK"
    if flavor == "J": r["n"] = "NK1"  # ReadString() accepts null as a key
    rows["OBJ_KEY"] = r
    r = token_row("OBJX_FIRST"); r["QUOTE"] = "KEY_BEGIN"; r["RBRACE"] = "POP"; rows["OBJX_FIRST"] = r
    r = token_row("OBJX_KEY"); r["QUOTE"] = "KEY_BEGIN"
    if flavor == "J": r["n"] = "NKX1"
    rows["OBJX_KEY"] = r
    r = token_row("STRUCT_FIRST"); r["QUOTE"] = "KEY_BEGIN"; r["RBRACE"] = "POP"; rows["STRUCT_FIRST"] = r
    r = token_row("STRUCT_KEY"); r["QUOTE"] = "KEY_BEGIN"; rows["STRUCT_KEY"] = r
    for c in "OSX":
        r = token_row("COLON_" + c); r["COLON"] = "VAL_" + c; rows["COLON_" + c] = r
    for c in "AOSX":
        rows[after[c]] = after_row(c)
    rows["FINISH"] = after_row("T")
    rows["STOP"] = {k: "STOP" for k in CLASSES}
    rows["ERRSTATE"] = {k: "ERRSTATE" for k in CLASSES}

    # strings
    for f in fams:
        key = not f.startswith("V_")
        for esc in (False, True):
            s = "STR" + f + ("_E" if esc else "")
            r = {}
            for k in CLASSES:
                if k == "QUOTE":
                    if f == "K": r[k] = "COLON_O"
                    elif f in ("S", "X"): r[k] = "KEY_END"
                    else:
                        c = f[2]
                        r[k] = "STR_DONE" if c in ("S", "X") else after[c]
                elif k == "BSLASH":
                    r[k] = "ESC" + f
                elif k in CTRLS:
                    if f == "S": r[k] = s  # readFieldHash never checks control characters
                    else: r[k] = s if (flavor == "J" and esc) else "ERR"  # jsoniter: only before the first backslash
                else:
                    r[k] = s
            rows[s] = r
        r = {k: "ERR" for k in CLASSES}
        r["u"] = "U4" + f
        for k in ESC_OK: r[k] = "STR" + f + "_E"
        rows["ESC" + f] = r
        for n in (4, 3, 2, 1):
            r = {k: "ERR" for k in CLASSES}
            for k in HEX: r[k] = ("U%d%s" % (n - 1, f)) if n > 1 else "STR" + f + "_E"
            rows["U%d%s" % (n, f)] = r
    # literals
    for c in ctxs:
        for l, (want, follow) in lit_chain.items():
            r = {k: "ERR" for k in CLASSES}
            r[want] = (follow + "_" + c) if follow else after[c]
            rows[l + "_" + c] = r
    for pre, colon in (("NK", "COLON_O"), ("NKX", "COLON_X")):
        for i, want in ((1, "u"), (2, "l"), (3, "l")):
            r = {k: "ERR" for k in CLASSES}
            r[want] = (pre + str(i + 1)) if i < 3 else colon
            rows[pre + str(i)] = r
    # numbers: BFS per context; the end of a number is folded into the after-value transition of its context, except in
    # context X where the engine must finish a captured counter first
    done = set()
    while work:
        st, c, n = work.pop()
        if n in done: continue
        done.add(n)
        r = {}
        for k in CLASSES:
            res = num_step(st, k)
            if res[0] == "err": r[k] = "ERR"
            elif res[0] == "done_re":
                r[k] = "NUM_DONE" if c == "X" else rows[after[c]][k]
            else: r[k] = num(res[1], c)
        rows[n] = r
    assert len(names) < EV_BASE, len(names)
    table = []
    for n in names:
        row = rows[n]
        ent = []
        for k in CLASSES:
            tgt = row[k]
            if tgt == "ERR": ent.append(EV["ERR"])
            elif tgt in EV: ent.append(EV[tgt])
            else: ent.append(sid[tgt])
        table.append(ent)
    return names, sid, table, n_str


def render():
    out = ["// GENERATED by tools/gen_json_tables.py — do not edit. Transition tables of the gateway's JSON automaton.",
           "// t = T[state * 32 + class]: t < 240 is the next state, t >= 240 an event (EV_*) handled by the engine.",
           "#pragma once", "#include <stdint.h>", "namespace arks {", f"static constexpr int kJsonClasses = {NCLS};"]
    out.append("#define ARKS_JSON_CLASS_TABLE {" + ",".join(str(CI[cls_of(b)]) for b in range(256)) + "}")
    want = ["STRS", "STRX", "COLON_S", "COLON_X", "OBJ_FIRST", "ARR_FIRST", "ARRC_FIRST", "OBJX_FIRST", "STRUCT_FIRST", "TOP",
            "AFTER_A", "AFTER_O", "AFTER_S", "AFTER_X", "FINISH", "STOP", "ERRSTATE", "VAL_A", "VAL_S", "VAL_X", "VAL_T",
            "VALG_S", "VALG_X", "VALG_T", "STRV_S", "STRV_X", "STRV_T"]
    ids = None
    for flavor in "JE":
        names, sid, table, n_str = build(flavor)
        flat = ",".join(str(x) for row in table for x in row)
        out.append(f"static constexpr int kJsonStates{flavor} = {len(names)};")
        out.append(f"#define ARKS_JSON_TABLE_{flavor} {{{flat}}}")
        cur = {n: sid[n] for n in want}
        cur["N_STRING_STATES"] = n_str
        if ids is None:
            ids = cur
        else:
            assert ids == cur, "the engine relies on these state ids being flavor independent"
    for n, v in ids.items():
        out.append(f"static constexpr uint32_t TS_{n} = {v};")
    out.append(f"static constexpr uint32_t EV_BASE = {EV_BASE};")
    out.append("enum : uint32_t { " + ", ".join(f"EV_{n} = {v}" for n, v in EV.items()) + " };")
    out.append("}  // namespace arks")
    return "\n".join(out) + "\n"


def path():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "arks_b200", "csrc", "json_tables.h")


if __name__ == "__main__":
    open(path(), "w").write(render())
    print("wrote", path(), "states J", len(build("J")[0]), "E", len(build("E")[0]))
