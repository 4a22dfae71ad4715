"""The `toss` import name (reference toss/__init__.py:1-49) and the staging of the reference's test files."""
import ast
import filecmp
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def test_toss_exports_every_name_of_the_reference_package():
    import toss
    import toss_b200
    names = ["FitnessFunctions", "estimate_covered_volume", "_compute_squared_distance", "compute_space_coverage",
             "update_spherical_tensor_grid", "create_spherical_tensor_grid", "cart2sphere", "sphere2cart", "get_fitness",
             "compute_trajectory", "compute_motion", "setup_spin_axis", "rotate_point", "get_trajectory_fixed_step",
             "get_trajectory_adaptive_step", "target_altitude_distance", "close_distance_penalty", "far_distance_penalty",
             "covered_volume", "total_covered_volume", "covered_space", "create_mesh", "is_outside", "IntegrationScheme",
             "plot_UDP_3D", "plot_UDP_2D", "fitness_over_generations", "fitness_over_time", "distance_deviation_over_time",
             "load_default_cfg", "udp_initial_condition", "setup_parameters", "setup_initial_state_domain"]
    if REF.exists():   # the list above IS the reference's __all__
        tree = ast.parse((REF / "toss" / "__init__.py").read_text())
        ref_all = next(ast.literal_eval(n.value) for n in tree.body
                       if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "__all__")
        assert sorted(ref_all) == sorted(names)
    for n in names:
        assert hasattr(toss, n), n
    assert sorted(toss.__all__) == sorted(names)
    # one implementation under two names: same function objects, same module objects (one device-context cache)
    assert toss.compute_trajectory is toss_b200.compute_trajectory
    from toss.trajectory.compute_trajectory import compute_trajectory
    from toss.optimization.udp_initial_condition import udp_initial_condition
    assert compute_trajectory is toss_b200.compute_trajectory and udp_initial_condition is toss_b200.udp_initial_condition
    assert sys.modules["toss._runtime"] is sys.modules["toss_b200._runtime"]
    with pytest.raises(NotImplementedError):
        toss.plot_UDP_3D()


def test_dotmap_shim_serves_the_reference_import():
    sys.path.append(str(ROOT / "toss" / "_shims"))
    try:
        from dotmap import DotMap
        d = DotMap({"a": {"b": 1}}, _dynamic=False)
        assert d.a.b == 1
    finally:
        sys.path.remove(str(ROOT / "toss" / "_shims"))


@pytest.mark.skipif(not REF.exists(), reason="build container only")
def test_staged_reference_tests_are_byte_identical():
    import __graft_entry__ as g
    g.stage_reference_tests()
    staged = ROOT / "baseline" / "_ref" / "tests"
    files = sorted(p.name for p in (REF / "tests").glob("*_test.py"))
    assert len(files) == 6
    for f in files:
        assert filecmp.cmp(REF / "tests" / f, staged / f, shallow=False), f
