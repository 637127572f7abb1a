#!/usr/bin/env python
"""Generates arks_b200/csrc/skip_dfa_tables.h: the table-driven automaton that validates SKIPPED JSON subtrees.

Inside a subtree whose content is irrelevant (anything below a key the gateway does not read) the only job is the
grammar check jsoniter's strict Skip() / encoding/json's checkValid would do. That is a pure pushdown automaton:
state x byte-class -> (next state, stack action). The per-byte step is two table lookups and no data-dependent
branching, so 32 lanes parsing 32 different documents stay converged.

The number states are not hand-written: they are discovered by breadth-first search over an executable copy of
JsonM::step_number's logic (trySkipNumber x RFC 8259 DFA product), so the table cannot drift from the reference
semantics restated in oracle/ork_json.c (and the CPU differential fuzz would catch it if it did).

Usage: python tools/gen_skip_dfa.py  (rewrites the header in place)
"""
import os

# ---- byte classes
CLASSES = ["OTHER", "SP", "WSC", "CTRL", "QUOTE", "BSLASH", "COMMA", "COLON", "LBRACE", "RBRACE", "LBRACK", "RBRACK",
           "MINUS", "PLUS", "DOT", "ZERO", "DIG19", "e", "E", "n", "t", "f", "u", "l", "r", "a", "s", "b", "SLASH",
           "HEXLO", "HEXUP"]
CI = {n: i for i, n in enumerate(CLASSES)}
NCLS = 32


def cls_of(b):
    c = chr(b)
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
ESC_OK = {"QUOTE", "BSLASH", "SLASH", "b", "f", "n", "r", "t"}

# ---- flags (table entry = next | flag << 6)
F_NONE, F_ERR, F_PUSHO, F_PUSHA, F_POPO, F_POPA, F_DONE, F_DONE_RE = range(8)

# fixed states; the four skippable string states come first (can_fast() == ss < 4)
FIXED = ["STRV", "STRV_E", "STRK", "STRK_E",
         "VAL", "ARR_FIRST", "OBJ_FIRST", "OBJ_KEY", "COLON", "AFTER_OBJ", "AFTER_ARR",
         "ESCV", "U4V", "U3V", "U2V", "U1V", "ESCK", "U4K", "U3K", "U2K", "U1K",
         "N1", "N2", "N3", "T1", "T2", "T3", "F1", "F2", "F3", "F4", "NK1", "NK2", "NK3"]

# ---- executable copy of JsonM::step_number on abstract state (tsn, any, dot, need, nf)
FM, FZ, FI, FD, FF, FE, FS, FX, DEAD = range(9)
ACCEPT = {FZ, FI, FF, FX}
TERM = {"COMMA", "RBRACK", "RBRACE", "SP", "WSC"}
NUMBYTE = DIGIT | {"DOT", "e", "E", "PLUS", "MINUS"}


def num_step(state, k):
    """-> ("err",) | ("done_re",) | ("next", state')"""
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
    if not tsn: any_, dot, need = 0, 0, 0  # irrelevant once trySkipNumber has given up: canonicalise
    return ("next", (tsn, any_, dot, need, nx))


def num_start(flavor, k):
    tsn = 1 if (flavor == "J" and k != "ZERO") else 0
    nf = FM if k == "MINUS" else FZ if k == "ZERO" else FI
    return (tsn, 0, 0, 0, nf)


def build(flavor):
    names = list(FIXED)
    num_id = {}
    work = []

    def nid(st):
        if st not in num_id:
            num_id[st] = len(names)
            names.append("NUM_%d%d%d%d_%d" % st)
            work.append(st)
        return num_id[st]

    rows = {}

    def value_start(k):
        if k == "QUOTE": return ("STRV", F_NONE)
        if k == "n": return ("N1", F_NONE)
        if k == "t": return ("T1", F_NONE)
        if k == "f": return ("F1", F_NONE)
        if k in ("MINUS", "ZERO", "DIG19"): return (nid(num_start(flavor, k)), F_NONE)
        if k == "LBRACK": return ("ARR_FIRST", F_PUSHA)
        if k == "LBRACE": return ("OBJ_FIRST", F_PUSHO)
        return ("VAL", F_ERR)

    lit_chain = {"N1": ("u", "N2"), "N2": ("l", "N3"), "N3": ("l", None), "T1": ("r", "T2"), "T2": ("u", "T3"),
                 "T3": ("e", None), "F1": ("a", "F2"), "F2": ("l", "F3"), "F3": ("s", "F4"), "F4": ("e", None),
                 "NK1": ("u", "NK2"), "NK2": ("l", "NK3"), "NK3": ("l", "KEYDONE")}
    for s in FIXED:
        row = []
        for k in CLASSES:
            nxt, fl = s, F_ERR
            if s in ("VAL", "ARR_FIRST"):
                if k in WS: nxt, fl = s, F_NONE
                elif s == "ARR_FIRST" and k == "RBRACK": nxt, fl = s, F_POPA
                else: nxt, fl = value_start(k)
            elif s == "OBJ_FIRST":
                if k in WS: fl = F_NONE
                elif k == "QUOTE": nxt, fl = "STRK", F_NONE
                elif k == "RBRACE": fl = F_POPO
            elif s == "OBJ_KEY":
                if k in WS: fl = F_NONE
                elif k == "QUOTE": nxt, fl = "STRK", F_NONE
                elif k == "n" and flavor == "J": nxt, fl = "NK1", F_NONE  # ReadString() accepts null as a key
            elif s == "COLON":
                if k in WS: fl = F_NONE
                elif k == "COLON": nxt, fl = "VAL", F_NONE
            elif s == "AFTER_OBJ":
                if k in WS: fl = F_NONE
                elif k == "COMMA": nxt, fl = "OBJ_KEY", F_NONE
                elif k == "RBRACE": fl = F_POPO
            elif s == "AFTER_ARR":
                if k in WS: fl = F_NONE
                elif k == "COMMA": nxt, fl = "VAL", F_NONE
                elif k == "RBRACK": fl = F_POPA
            elif s in ("STRV", "STRV_E", "STRK", "STRK_E"):
                key, esc = s.startswith("STRK"), s.endswith("_E")
                if k == "QUOTE": nxt, fl = ("COLON", F_NONE) if key else (s, F_DONE)
                elif k == "BSLASH": nxt, fl = ("ESCK" if key else "ESCV"), F_NONE
                elif k in ("WSC", "CTRL"):
                    # jsoniter: control characters are only rejected before the string's first backslash; encoding/json: always
                    fl = F_NONE if (flavor == "J" and esc) else F_ERR
                else: fl = F_NONE
            elif s in ("ESCV", "ESCK"):
                v = "V" if s == "ESCV" else "K"
                if k == "u": nxt, fl = "U4" + v, F_NONE
                elif k in ESC_OK: nxt, fl = ("STRV_E" if v == "V" else "STRK_E"), F_NONE
            elif s[0] == "U" and s[1] in "4321":
                v, n = s[2], int(s[1])
                if k in HEX: nxt, fl = (("U%d%s" % (n - 1, v)) if n > 1 else ("STRV_E" if v == "V" else "STRK_E")), F_NONE
            elif s in lit_chain:
                want, follow = lit_chain[s]
                if k == want:
                    if follow is None: nxt, fl = s, F_DONE
                    elif follow == "KEYDONE": nxt, fl = "COLON", F_NONE
                    else: nxt, fl = follow, F_NONE
            row.append((nxt, fl))
        rows[s] = row
    # numbers: BFS
    done = set()
    while work:
        st = work.pop()
        if st in done: continue
        done.add(st)
        row = []
        for k in CLASSES:
            r = num_step(st, k)
            if r[0] == "err": row.append((nid(st), F_ERR))
            elif r[0] == "done_re": row.append((nid(st), F_DONE_RE))
            else: row.append((nid(r[1]), F_NONE))
        rows[num_id[st]] = row
    sid = {n: i for i, n in enumerate(names)}
    assert len(names) <= 64, len(names)
    table = []
    for i, n in enumerate(names):
        row = rows[n] if n in rows else rows[i]
        ent = []
        for nxt, fl in row:
            j = nxt if isinstance(nxt, int) else sid[nxt]
            ent.append(j | (fl << 6))
        ent += [sid.get("VAL", 4) | (F_ERR << 6)] * (NCLS - len(ent))
        table.append(ent)
    return names, table


def main():
    out = ["// GENERATED by tools/gen_skip_dfa.py — do not edit. Table-driven automaton for skipped JSON subtrees.",
           "// entry = next_state | flag << 6; flags: 0 none 1 error 2 push-object 3 push-array 4 pop-object 5 pop-array",
           "//         6 scalar done (byte consumed) 7 number done before this byte (re-dispatch the byte).",
           "#pragma once", "#include <stdint.h>", "namespace arks {", f"static constexpr int kSkipClasses = {NCLS};"]
    out.append("#define ARKS_SKIP_CLASS_TABLE {" + ",".join(str(CI[cls_of(b)]) for b in range(256)) + "}")
    n_states = {}
    for flavor in "JE":
        names, table = build(flavor)
        n_states[flavor] = len(names)
        flat = ",".join(str(x) for row in table for x in row)
        out.append(f"static constexpr int kSkipStates{flavor} = {len(names)};")
        out.append(f"#define ARKS_SKIP_TABLE_{flavor} {{{flat}}}")
        if flavor == "J":
            for n in ("STRV", "STRV_E", "STRK", "STRK_E", "ARR_FIRST", "OBJ_FIRST", "AFTER_OBJ", "AFTER_ARR"):
                out.append(f"static constexpr uint32_t K_{n} = {names.index(n)};")
    out.append("enum : uint32_t { SKF_NONE = 0, SKF_ERR, SKF_PUSHO, SKF_PUSHA, SKF_POPO, SKF_POPA, SKF_DONE, SKF_DONE_RE };")
    out.append("}  // namespace arks")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "arks_b200", "csrc", "skip_dfa_tables.h")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path, "states", n_states)


if __name__ == "__main__":
    main()
