"""Design prototypes kept runnable: algorithm statements the CUDA kernels (or their planned successors) follow."""
import importlib.util
import os

from conftest import ROOT


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", "bringup", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_chain_pass_as_scans_matches_the_serial_pass():
    # round-2 groundwork: the decoder's serial chain pass restated as scans + one sort by stack level
    _load("proto_chain_scan").main()


def test_prefix_code_parse_as_a_scan_matches_the_serial_parse():
    # round-2 groundwork: the order-N bitstream's code boundaries from a three-state automaton scan
    _load("proto_markov_parse").main()
