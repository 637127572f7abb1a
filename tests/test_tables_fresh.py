"""The automaton tables (arks_b200/csrc/json_tables.h) are generated: the committed header must be what the generator emits."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_json_tables_header_is_up_to_date():
    spec = importlib.util.spec_from_file_location("gen_json_tables", os.path.join(ROOT, "tools", "gen_json_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert open(gen.path()).read() == gen.render(), "run `python tools/gen_json_tables.py`"
    # sanity of the layout the engine relies on: string states first and in (plain, escaped) pairs, events >= 240
    names, sid, table, n_str = gen.build("J")
    assert n_str % 2 == 0 and all(names[i].startswith("STR") for i in range(n_str))
    assert all(names[i + 1] == names[i] + "_E" for i in range(0, n_str, 2))
    assert len(names) < gen.EV_BASE and max(max(r) for r in table) < 256
