"""params.py surface parity with the reference (reference src/2d/params.py:13-107)."""
import json

import pytest

from heat_pipe_solver_b200 import params


@pytest.fixture(autouse=True)
def _fresh():
    params.reset_cache()
    yield
    params.reset_cache()


def test_flat_and_groups_match_reference(golden_params):
    assert params.get_all_params() == golden_params['flat']
    assert params.load_params() == golden_params['flat']
    for g, body in golden_params['groups'].items():
        assert params.get_param_group(g) == body
    assert params.get_param_group('nope') == golden_params['unknown_group'] == {}
    assert params.get_param_group(None) == golden_params['none_group'] == {}


def test_get_param_default():
    assert params.get_param('T_amb') == 290.0
    assert params.get_param('does_not_exist') is None
    assert params.get_param('does_not_exist', 3) == 3


def test_cache_ignores_later_config_file(tmp_path):
    # reference params.py:26-27: once cached, config_file is ignored
    first = params.load_params()
    other = tmp_path / 'other.json'
    other.write_text(json.dumps({'parameters': {'T_amb': 1.0}}))
    assert params.load_params(str(other)) is first
    params.reset_cache()
    assert params.load_params(str(other)) == {'T_amb': 1.0}
    # get_param_group always re-reads the DEFAULT file (reference params.py:84-87)
    assert params.get_param_group('parameters')['T_amb'] == 290.0


def test_errors(tmp_path):
    with pytest.raises(FileNotFoundError, match='Configuration file not found'):
        params.load_params(str(tmp_path / 'missing.json'))
    bad = tmp_path / 'bad.json'
    bad.write_text('{not json')
    with pytest.raises(ValueError, match='Invalid JSON in configuration file'):
        params.load_params(str(bad))
