"""Host Config mirror vs the reference's Config::load / compute_derived (src/config.cpp)."""
import os

import pytest

import helpers as H
from oracle import refapi
from pd_mg_pin_corrosion_b200.config import Config, PdConfig


def test_defaults_and_derived():
    c = Config.load(None, quiet=True)
    assert c.dx == 5.0e-6 and c.m_ratio == 3 and c.use_implicit == 1      # src/config.h:6-7,72
    assert c.delta == 3 * 5.0e-6
    assert c.U_in == c.Q_flow / (3.14159265358979323846 * c.R_tube * c.R_tube)
    assert c.c0 == 25.0 * c.U_in                                          # src/config.cpp:107-110


def test_missing_file_warns_and_uses_defaults(capsys):
    c = Config.load("/nonexistent/params.cfg", quiet=True)
    assert "Cannot open config file" in capsys.readouterr().err           # src/config.cpp:18-23
    assert c.R_tube == 150.0e-6


def test_parser_semantics(tmp_path, capsys):
    p = tmp_path / "a.cfg"
    p.write_text("# comment\n dx = 2.0e-6   # trailing\nm_ratio=3\nbogus_key = 1\nflow_max_iters = 12abc\n"
                 "output_dir = out dir\nL_wire =\n= 5\ndx = 4.0e-6\n")
    c = Config.load(str(p), quiet=True)
    assert "Unknown config key 'bogus_key'" in capsys.readouterr().err
    assert c.dx == 4.0e-6                       # later key wins (sequential scan)
    assert c.flow_max_iters == 12               # std::stoi prefix semantics
    assert c.output_dir == "out dir"
    assert c.L_wire == 400.0e-6                 # empty value ignored


def test_struct_roundtrip():
    _, c, _ = H.load_cfg("2d_default")
    s = c.to_struct()
    assert isinstance(s, PdConfig)
    for name, _ in PdConfig._fields_:
        if name != "reserved":
            assert getattr(s, name) == getattr(c, name), name


def test_out_of_scope_branches_are_rejected():
    c = Config.load(os.path.join(H.CONFIG_DIR, "params.cfg"), quiet=True)   # as shipped: implicit (built)
    c.check_supported()
    c = Config.load(None, {"use_amr": 1}, quiet=True)
    with pytest.raises(ValueError, match="use_amr"):
        c.check_supported()


@pytest.mark.skipif(not refapi.have_ref(2), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["2d_default", "2d_poiseuille", "2d_offgrid", "2d_dissolve"])
def test_matches_reference_parser(case):
    dim, cfg, ov = H.load_cfg(case)
    base = H.CASES[case][1]
    r = refapi.RefSim(dim, base, ov, build=False)
    for k, v in r.cfg.items():
        assert getattr(cfg, k) == v, k
