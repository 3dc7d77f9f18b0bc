"""Python-level heuristics written against the QRMSAEnv call surface (single-env compatibility path).

Same names, arguments and return convention `(action, blocked_due_to_resources, blocked_due_to_osnr)` as
reference heuristics/heuristics.py:36-54 and :923-966.  The batched fast path does not go through here:
`BatchedQRMSAEnv.step_first_fit` runs the same policy fused with the step inside one CUDA kernel; this
module exists so that user-written Python heuristics (which read `get_available_slots`, `_get_candidates`,
`get_number_slots`, `calculate_osnr`) keep working when the env is switched to the B200 one.
"""
from __future__ import annotations

from .env import QRMSAEnv, calculate_osnr


def get_qrmsa_env(env) -> QRMSAEnv:
    """Unwrap `.env` chains down to the QRMSAEnv (heuristics.py:15-33)."""
    while not isinstance(env, QRMSAEnv):
        if not hasattr(env, "env"):
            raise ValueError("no QRMSAEnv found in the wrapper chain")
        env = env.env
    return env


def get_action_index(env: QRMSAEnv, path_index: int, modulation_index: int, initial_slot: int) -> int:
    """(path, absolute modulation index, slot) -> action (heuristics.py:36-54)."""
    relative = env.max_modulation_idx - modulation_index
    return (path_index * env.modulations_to_consider * env.num_spectrum_resources
            + relative * env.num_spectrum_resources + initial_slot)


def heuristic_shortest_available_path_first_fit_best_modulation(env):
    """Shortest path first, most efficient modulation first, first-fit slot, accept on GSNR >= threshold
    (heuristics.py:923-966)."""
    sim = get_qrmsa_env(env)
    svc = sim.current_service
    blocked_resources = blocked_osnr = False
    for path_idx, path in enumerate(sim.k_shortest_paths[svc.source, svc.destination]):
        available = sim.get_available_slots(path)
        for modulation_idx in range(sim.max_modulation_idx, -1, -1):
            modulation = sim.modulations[modulation_idx]
            n = sim.get_number_slots(svc, modulation)
            if n <= 0:
                continue
            starts = sim._get_candidates(available, n, sim.num_spectrum_resources)
            if not starts:
                blocked_resources = True
                continue
            svc.path, svc.initial_slot, svc.number_slots, svc.current_modulation = path, starts[0], n, modulation
            svc.center_frequency = (sim.frequency_start + (sim.frequency_slot_bandwidth * starts[0])
                                    + (sim.frequency_slot_bandwidth * (n / 2)))
            svc.bandwidth = sim.frequency_slot_bandwidth * n
            svc.launch_power = sim.launch_power
            osnr, _, _ = calculate_osnr(sim, svc)
            if osnr >= modulation.minimum_osnr + sim.margin:
                return get_action_index(sim, path_idx, modulation_idx, starts[0]), False, False
            blocked_osnr = True
            blocked_resources = False
    return sim.action_space.n - 1, blocked_resources, blocked_osnr


# identical decisions in the reference (heuristics.py:431-490)
shortest_available_path_lowest_spectrum_best_modulation = heuristic_shortest_available_path_first_fit_best_modulation


def load_balancing_best_modulation(env):
    """Least-loaded path (occupied slots of the path availability / hops) that admits a modulation; best
    modulation and first-fit slot on it (heuristics.py:547-627).  Device counterpart: policy "load_balancing"."""
    import numpy as np

    sim = get_qrmsa_env(env)
    svc = sim.current_service
    solution, lowest = None, float("inf")
    any_res = any_osnr = False
    for path_idx, path in enumerate(sim.k_shortest_paths[svc.source, svc.destination]):
        available = sim.get_available_slots(path)
        load = np.sum(available == 0) / len(path.links)
        if load >= lowest:
            continue
        for modulation_idx in range(sim.max_modulation_idx, -1, -1):
            modulation = sim.modulations[modulation_idx]
            n = sim.get_number_slots(svc, modulation)
            if n <= 0:
                continue
            starts = sim._get_candidates(available, n, sim.num_spectrum_resources)
            if not starts:
                any_res = True
                continue
            svc.path, svc.initial_slot, svc.number_slots, svc.current_modulation = path, starts[0], n, modulation
            osnr, _, _ = calculate_osnr(sim, svc)
            if osnr >= modulation.minimum_osnr + sim.margin:
                lowest = load
                solution = get_action_index(sim, path_idx, modulation_idx, starts[0])
                break
            any_osnr = True
    if solution is not None:
        return solution, False, False
    if any_osnr:
        any_res = False
    return sim.action_space.n - 1, any_res, any_osnr


def heuristic_highest_snr(env):
    """Every valid start of every (path, modulation) is QoT-checked; the acceptable candidate with the highest GSNR
    wins, the first one met keeps a tie (heuristics.py:272-328, the benchmark's heuristic #2).  This Python form makes
    one device probe per candidate (thousands per request); the batched path runs the same search fused in one kernel:
    `BatchedQRMSAEnv.step_heuristic("highest_snr", n)` / `qrmsa_step_heuristic(QRMSA_POLICY_HIGHEST_SNR)`."""
    sim = get_qrmsa_env(env)
    svc = sim.current_service
    best_osnr, best_action = float("-inf"), None
    any_res = any_osnr = False
    for path_idx, path in enumerate(sim.k_shortest_paths[svc.source, svc.destination]):
        for modulation_idx in range(sim.max_modulation_idx, -1, -1):
            modulation = sim.modulations[modulation_idx]
            n = sim.get_number_slots(svc, modulation)
            if n <= 0:
                continue
            starts = sim._get_candidates(sim.get_available_slots(path), n, sim.num_spectrum_resources)
            if not starts:
                any_res = True
                continue
            for start in starts:
                svc.path, svc.initial_slot, svc.number_slots, svc.current_modulation = path, start, n, modulation
                svc.center_frequency = (sim.frequency_start + (sim.frequency_slot_bandwidth * start)
                                        + (sim.frequency_slot_bandwidth * (n / 2)))
                svc.bandwidth = sim.frequency_slot_bandwidth * n
                svc.launch_power = sim.launch_power
                osnr, _, _ = calculate_osnr(sim, svc)
                if osnr >= modulation.minimum_osnr + sim.margin:
                    if osnr > best_osnr:
                        best_osnr, best_action = osnr, get_action_index(sim, path_idx, modulation_idx, start)
                else:
                    any_osnr = True
    if best_action is not None:
        return best_action, False, False
    if any_osnr:
        any_res = False
    return sim.action_space.n - 1, any_res, any_osnr


def heuristic_load_balancing_first_fit(env):
    """The k paths ordered by the occupied fraction of their availability (ties: path index), then best modulation and
    first-fit slot on the first path that admits one; a reject reports (True, False) (heuristics.py:202-270).  Device
    counterpart: policy "load_balancing_first_fit"."""
    import numpy as np

    sim = get_qrmsa_env(env)
    svc = sim.current_service
    order = []
    for path_idx, path in enumerate(sim.k_shortest_paths[svc.source, svc.destination]):
        available = sim.get_available_slots(path)
        order.append((np.sum(available == 0) / len(available) if len(available) > 0 else 1.0, path_idx, path))
    order.sort(key=lambda t: (t[0], t[1]))
    for _, path_idx, path in order:
        for modulation_idx in range(sim.max_modulation_idx, -1, -1):
            modulation = sim.modulations[modulation_idx]
            n = sim.get_number_slots(svc, modulation)
            if n <= 0:
                continue
            starts = sim._get_candidates(sim.get_available_slots(path), n, sim.num_spectrum_resources)
            if not starts:
                continue
            svc.path, svc.initial_slot, svc.number_slots, svc.current_modulation = path, starts[0], n, modulation
            svc.center_frequency = (sim.frequency_start + (sim.frequency_slot_bandwidth * starts[0])
                                    + (sim.frequency_slot_bandwidth * (n / 2)))
            svc.bandwidth = sim.frequency_slot_bandwidth * n
            svc.launch_power = sim.launch_power
            osnr, _, _ = calculate_osnr(sim, svc)
            if osnr >= modulation.minimum_osnr + sim.margin:
                return get_action_index(sim, path_idx, modulation_idx, starts[0]), False, False
    return sim.action_space.n - 1, True, False
