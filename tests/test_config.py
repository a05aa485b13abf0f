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


SHIPPED = ["params.cfg", "params_amr.cfg", "params_amr_r2.cfg", "params_calibration.cfg", "params_calibration_v2.cfg",
           "params_diagnostic.cfg", "params_fine.cfg", "params_fine_calibration.cfg", "params_implicit_test.cfg",
           "params_poiseuille.cfg", "params_transport_viz.cfg"]


@pytest.mark.skipif(not refapi.have_ref(2), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", SHIPPED)
def test_every_shipped_config_parses_like_the_reference(name):
    """configs/<name> (the parameter values of every configuration file the reference ships) through the Python
    mirror against the reference's own Config::load + compute_derived on the same file: every numeric member incl. the
    implicit / AMR keys and the derived values; and, where the reference tree is present, configs/<name> against the
    reference's own config/<name>."""
    path = os.path.join(H.CONFIG_DIR, name)
    want = refapi.parse_config_with_reference(2, path)
    cfg = Config.load(path, quiet=True)
    for k, v in want.items():
        assert getattr(cfg, k) == v, (name, k, getattr(cfg, k), v)
    theirs = os.path.join("/root/reference/config", name)
    if os.path.exists(theirs):
        assert refapi.parse_config_with_reference(2, theirs) == want, name


@pytest.mark.skipif(not refapi.have_ref(2), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", SHIPPED)
def test_cpp_driver_parser_matches_the_reference(name):
    """host/config.cpp (what host/pd_corrosion_gpu reads its configuration with) against the reference's Config::load
    on every shipped configuration: the PdConfig that crosses the C ABI and the members that stay on the host."""
    import ctypes as C
    from pd_mg_pin_corrosion_b200 import grains as G
    L = G._load()
    L.pdhost_load_config.restype = C.c_int
    L.pdhost_load_config.argtypes = [C.c_char_p, C.POINTER(PdConfig), C.POINTER(C.c_double), C.c_char_p, C.c_int]
    path = os.path.join(H.CONFIG_DIR, name)
    pod, extra, out = PdConfig(), (C.c_double * 14)(), C.create_string_buffer(512)
    assert L.pdhost_load_config(path.encode(), C.byref(pod), extra, out, 512) == 0
    want = refapi.parse_config_with_reference(2, path)
    for k, _ in PdConfig._fields_:
        if k != "reserved":
            assert getattr(pod, k) == want[k], (name, k)
    host_members = ["implicit_dt_fraction", "implicit_dt_max", "implicit_output_every", "diagnostic_every", "newton_tol",
                    "newton_max_iter", "use_amr", "amr_ratio", "amr_buffer", "precip_fraction", "grain_size_mean",
                    "gb_width_cells", "precip_cluster_cells", "C_sat"]
    for k, v in zip(host_members, list(extra)):
        assert v == want[k], (name, k)
    assert out.value.decode() == want["output_dir"]
