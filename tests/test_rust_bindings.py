"""The Rust boundary ships as files (rust/zaru-b200-sys, rust/zaru-b200) but cannot be compiled in this image (no
cargo / rustc).  What CAN be checked on the CPU: the generated `extern "C"` block declares exactly the symbols of
include/zaru_b200.h with the same number of arguments and the translated types, it is up to date with the generator,
and every `sys::zb_*` call in the adapter names a declared symbol with the right arity."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import gen_rust_sys  # noqa: E402


def _rust_externs():
    text = open(os.path.join(ROOT, "rust", "zaru-b200-sys", "src", "lib.rs")).read()
    block = text[text.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    out = {}
    for m in re.finditer(r"pub fn (zb_[a-z0-9_]+)\((.*?)\)(?: -> ([^;]+))?;", block):
        args = [a.strip() for a in m.group(2).split(", ") if a.strip()]
        out[m.group(1)] = ([a.split(": ", 1)[1] for a in args], (m.group(3) or "").strip())
    return out


def test_sys_crate_matches_the_header_symbol_for_symbol():
    decls = gen_rust_sys.declarations(open(gen_rust_sys.HEADER).read())
    rust = _rust_externs()
    assert sorted(rust) == sorted(name for _, name, _ in decls)
    assert len(rust) >= 70
    for ret, name, params in decls:
        rtypes, rret = rust[name]
        assert len(rtypes) == len(params), name
        assert rtypes == [gen_rust_sys.rust_type(t) for t, _ in params], name
        assert rret == ("" if ret == "void" else gen_rust_sys.rust_type(ret)), name
    # spot checks of the type translation itself
    assert rust["zb_ctx_create"][0] == ["i32", "*mut *mut zb_ctx"]
    assert rust["zb_net_estimate"][0] == ["*mut zb_net", "*const f32", "i32", "*const *mut f32"]
    assert rust["zb_net_input_info"][0][2] == "*mut *const c_char"
    assert rust["zb_last_error"] == ([], "*const c_char")
    assert rust["zb_frames_clear"][0][3] == "*const u8"


def test_sys_crate_is_up_to_date_with_its_generator():
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"]) == 0, \
        "include/zaru_b200.h changed: run python tools/gen_rust_sys.py"


def test_adapter_calls_only_declared_symbols_with_the_right_arity():
    rust = _rust_externs()
    text = open(os.path.join(ROOT, "rust", "zaru-b200", "src", "lib.rs")).read()
    calls = list(re.finditer(r"sys::(zb_[a-z0-9_]+)\(", text))
    assert len(calls) >= 45
    used = set()
    for m in calls:
        name = m.group(1)
        assert name in rust, f"adapter calls undeclared symbol {name}"
        depth, i, nargs, cur = 1, m.end(), 0, False
        while depth:                      # count top-level commas up to the matching parenthesis
            ch = text[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                nargs += 1
                cur = False
                i += 1
                continue
            if depth and not ch.isspace():
                cur = True
            i += 1
        nargs += 1 if cur else 0
        assert nargs == len(rust[name][0]), (name, nargs, len(rust[name][0]))
        used.add(name)
    # the entry points the round-1 review found missing from the Rust side
    for must in ("zb_frames_alias", "zb_frames_update", "zb_detector_extract", "zb_sync", "zb_hand_pipeline_run",
                 "zb_detector_timers", "zb_estimator_timers", "zb_face_iris_pipeline_run", "zb_tracker_track"):
        assert must in used, must
