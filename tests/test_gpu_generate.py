"""On-device request generation (qrmsa_generate_trace): the device stream against its numpy
restatement (oracle.generate_trace_philox), its statistics against the reference's traffic model
(qrmsa.pyx:1079-1099, :1124-1148), and decision parity of the fused step on the generated stream."""
import numpy as np
import pytest

from helpers import load_tables

pytestmark = pytest.mark.gpu


def _engine(tb, n_envs, n_req):
    from optical_networking_gym_b200.engine import Engine

    return Engine(tb, n_envs, n_req)


def test_device_stream_matches_numpy_restatement_and_continues():
    from optical_networking_gym_b200.tracegen import choice_tables
    from oracle import oracle as orc

    tb = load_tables("nobel-eu", 320)
    n_envs, n_req = 96, 150
    loads = np.linspace(100.0, 500.0, n_envs)
    eng = _engine(tb, n_envs, n_req)
    eng.reset()
    eng.generate_trace(n_req, loads, seed=0x1234567890AB, env_offset=1000)
    dev = eng.trace_host()
    tabs = choice_tables(tb.n_nodes, tb.n_rates)
    ref = orc.generate_trace_philox(n_envs, n_req, loads, 0x1234567890AB, *tabs, env_offset=1000)
    for k in range(3):
        assert np.array_equal(dev[k], ref[k]), ("src", "dst", "rate")[k]
    assert np.allclose(dev[3], ref[3], rtol=2e-7, atol=0) and np.allclose(dev[4], ref[4], rtol=2e-7, atol=0)
    assert (dev[0] != dev[1]).all()
    # a second call continues the streams and the clocks; a shard with the matching offset draws the same requests
    eng.reset()
    eng.generate_trace(50, loads, seed=0x1234567890AB, env_offset=1000, restart=False)
    nxt = eng.trace_host()
    ref2 = orc.generate_trace_philox(n_envs, n_req + 50, loads, 0x1234567890AB, *tabs, env_offset=1000)
    assert np.array_equal(nxt[0], ref2[0][n_req:]) and np.allclose(nxt[3], ref2[3][n_req:], rtol=2e-7)
    eng.close()
    shard = _engine(tb, 32, n_req)
    shard.reset()
    shard.generate_trace(n_req, loads[64:], seed=0x1234567890AB, env_offset=1064)
    sh = shard.trace_host()
    assert np.array_equal(sh[0], dev[0][:, 64:]) and np.array_equal(sh[3], dev[3][:, 64:])
    shard.close()


def test_traffic_statistics():
    tb = load_tables("nobel-eu", 320)
    n_envs, n_req, load = 2048, 400, 300.0
    eng = _engine(tb, n_envs, n_req)
    eng.reset()
    eng.generate_trace(n_req, load, seed=7)
    src, dst, rate, arrival, holding = eng.trace_host()
    iat = np.diff(np.vstack([np.zeros((1, n_envs), np.float64), arrival.astype(np.float64)]), axis=0)
    n = iat.size
    assert (iat >= 0).all()
    # exponential with mean holding/load (qrmsa.pyx:1130): sample mean within 5 sigma
    assert abs(iat.mean() - 10800.0 / load) < 5 * (10800.0 / load) / np.sqrt(n)
    assert abs(holding.mean() - 10800.0) < 5 * 10800.0 / np.sqrt(n)
    for arr, k in ((src, tb.n_nodes), (rate, tb.n_rates)):
        cnt = np.bincount(arr.ravel(), minlength=k)
        assert np.abs(cnt - n / k).max() < 6 * np.sqrt(n / k)
    assert (src != dst).all()
    eng.close()


def test_step_parity_on_generated_stream():
    """The fused step on a device-generated stream decides exactly as the oracle replaying the same stream."""
    from optical_networking_gym_b200 import _lib
    from optical_networking_gym_b200.engine import unpack_bitmaps
    from oracle import oracle as orc

    tb = load_tables("nobel-eu", 320)
    n_envs, n_req = 24, 501
    eng = _engine(tb, n_envs, n_req)
    eng.reset()
    eng.generate_trace(n_req, 350.0, seed=99)
    eng.step_first_fit(n_req - 1)
    tr = eng.trace_host()
    words = eng.actions_host(0, n_req - 1)
    actions = (words & _lib.ACTION_MASK).T
    flagged = ((words.view(np.uint32) & _lib.FLAG_NEAR_THRESHOLD) != 0).T
    slots = unpack_bitmaps(eng.export_bitmaps(0, n_envs), tb.n_slots)
    for e in range(n_envs):
        o = orc.OracleEnv(tb, n_req)
        o.reset(*[a[:, e] for a in tr])
        ref = o.run_first_fit(n_req - 1, log_qot=False)
        if not np.array_equal(ref["action"], actions[e]):
            d = int(np.flatnonzero(ref["action"] != actions[e])[0])
            assert flagged[e, d], f"env {e}: unflagged decision mismatch at step {d}"
            continue
        assert np.array_equal(o.slots(), slots[e])
    assert eng.counters_dict()["errors"] == 0
    eng.close()


def test_batched_env_with_device_traffic():
    """BatchedQRMSAEnv(request_source="device"): episodes continue the Philox streams, decisions match the oracle on the
    stream the env reports through current_requests()."""
    from optical_networking_gym_b200.env import BatchedQRMSAEnv
    from oracle import oracle as orc

    tb = load_tables("nsfnet", 320)
    n_envs, L = 17, 180
    env = BatchedQRMSAEnv(tb, n_envs, num_spectrum_resources=320, episode_length=L, load=300.0,
                          bit_rates=(10, 40, 100, 400, 1000), launch_power_dbm=1.0, seed=4242, request_source="device",
                          env_offset=5)
    first = [a.copy() for a in env.current_requests()]
    env.step_first_fit(L - 1)
    actions, flagged = env.actions(0, L - 1)
    actions, flagged = actions.T, flagged.T          # -> [env][step]
    for e in range(n_envs):
        o = orc.OracleEnv(tb, L)
        o.reset(*[a[:, e] for a in first])
        ref = o.run_first_fit(L - 1, log_qot=False)
        if not np.array_equal(ref["action"], actions[e]):
            d = int(np.flatnonzero(ref["action"] != actions[e])[0])
            assert flagged[e, d]
    env.reset()
    second = env.current_requests()
    assert (second[3][0] > first[3][-1]).all()        # the clocks run on across episodes (qrmsa.pyx:179, :1081)
    assert not np.array_equal(second[0], first[0])
    env.close()
