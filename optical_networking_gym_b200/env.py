"""QRMSAEnv-compatible host API over the CUDA engine.

Two classes:

* `QRMSAEnv`  -- ONE environment with the reference's constructor, `reset(seed, options)`,
  `step(action:int) -> (obs, reward, terminated, truncated, info)` and the attribute / method surface the
  reference's heuristics read (reference envs/qrmsa.pyx:206-237, :427-504, :838-1065; heuristics/
  heuristics.py:923-966).  State lives on the GPU (engine with n_envs = 1); helper methods pull it back.
  This is the compatibility path: correct, not fast.
* `BatchedQRMSAEnv` -- n_envs environments stepped together: `reset()`, `step(actions[n_envs])`,
  `step_first_fit(n_steps)` (the fused heuristic + step of the benchmark loop,
  examples/JOCN_Benchmark_2024/graph_load.py:161-163) and `action_masks()`.

Request streams: env i replays `random.Random(base_seed + i)` exactly as the reference draws it
(tracegen.py).  The reference seeds its generator from OS entropy (qrmsa.pyx:241) -- its `seed=` argument
does not reach it -- so any seed reproduces *a* valid reference run; passing the same seed the oracle
harness patches in reproduces *that* run bit for bit.

`gen_observation=True` returns the reference's observation vector and GSNR-validated action mask
(qrmsa.pyx:583-781) from the device kernel `qrmsa_observation`; with `gen_observation=False` both are zeros, as in
the reference (qrmsa.pyx:584-587).

`file_name=` writes the reference's per-service CSV (qrmsa.pyx:387-406, :967-990); `bit_rate_selection="continuous"`
draws `rng.randint(lower, higher)` like the reference (qrmsa.pyx:246-254); `reset(options={"only_episode_counters":
True})` restarts the episode counters on the live network and drops the pending release events (qrmsa.pyx:433, :461-464).

`measure_disruptions=True` and `defragmentation=True` / `n_defrag_services` run on the device (qrmsa.pyx:937-952, :1113-1122,
:1545-1639; `qrmsa_set_features`).

Not implemented: `bands` (multiband: SURVEY §8f ranks it last and documents why its reference path is unusable); it raises.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .engine import Engine, unpack_bitmaps
from .tables import StaticTables
from .tracegen import TraceGenerator


# --------------------------------------------------------------------------------------------------------
# tiny stand-ins for gymnasium.spaces (gymnasium is optional; only `.n` / `.shape` / `sample` are used by the
# reference's callers: heuristics.py `env.action_space.n - 1`, SB3 reads shape/dtype)
# --------------------------------------------------------------------------------------------------------
class Discrete:
    def __init__(self, n: int, seed=None):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return int(self._rng.integers(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class Service:
    """Request + allocation record, same field names as the reference Service (qrmsa.pyx:29-116)."""

    def __init__(self, service_id, source, source_id, destination, destination_id, arrival_time, holding_time,
                 bit_rate):
        self.service_id = int(service_id)
        self.source = source
        self.source_id = int(source_id)
        self.destination = destination
        self.destination_id = str(destination_id)      # a str in the reference (qrmsa.pyx:34,1146)
        self.arrival_time = float(np.float32(arrival_time))
        self.holding_time = float(np.float32(holding_time))
        self.bit_rate = float(np.float32(bit_rate))
        self.path = None
        self.initial_slot = -1
        self.number_slots = 0
        self.center_frequency = 0.0
        self.bandwidth = 0.0
        self.launch_power = 0.0
        self.current_modulation = None
        self.accepted = False
        self.blocked_due_to_resources = False
        self.blocked_due_to_osnr = False
        self.OSNR = self.ASE = self.NLI = 0.0

    def __repr__(self):
        return (f"Service(id={self.service_id}, {self.source}->{self.destination}, bit_rate={self.bit_rate}, "
                f"arrival={self.arrival_time}, holding={self.holding_time})")


def _unsupported(name):
    raise NotImplementedError(f"{name} is not implemented in the B200 path yet (SURVEY §8f 'next' rows)")


def _check_kwargs(measure_disruptions, defragmentation, bands, gen_observation, bit_rate_selection):
    if bands:
        _unsupported("bands (multiband)")
    if bit_rate_selection not in ("discrete", "continuous"):
        raise ValueError("bit_rate_selection must be 'discrete' or 'continuous'")


def _continuous_rates(lower, higher):
    """bit_rate_selection="continuous": rng.randint(int(lower), int(higher)) (qrmsa.pyx:246-254) -- integer rates, i.e.
    the table (lower, lower+1, ..., higher) indexed by (rate - lower)."""
    lo, hi = int(lower), int(higher)
    if lo < 1 or hi < lo or hi - lo + 1 > 255:
        raise ValueError("continuous bit rates: need 1 <= lower <= higher and at most 255 integer rates")
    return tuple(range(lo, hi + 1)), (lo, hi)


class _Common:
    """Shared construction: static tables from a reference topology graph (or a StaticTables)."""

    def _setup(self, topology, num_spectrum_resources, bit_rates, launch_power_dbm, margin, frequency_start,
               frequency_slot_bandwidth, channel_width, k_paths, modulations_to_consider, bandwidth):
        if isinstance(topology, StaticTables):
            self.topology = None
            self.tables = topology
        else:
            self.topology = topology
            self.tables = StaticTables.from_topology(
                topology, num_spectrum_resources=num_spectrum_resources, bit_rates=bit_rates,
                launch_power_dbm=launch_power_dbm, margin=margin, frequency_start=frequency_start,
                frequency_slot_bandwidth=frequency_slot_bandwidth, channel_width=channel_width, k_paths=k_paths,
                modulations_to_consider=modulations_to_consider)
        tb = self.tables
        frequency_end = frequency_start + frequency_slot_bandwidth * tb.n_slots
        assert math.isclose(frequency_end - frequency_start, bandwidth, rel_tol=1e-5), \
            "bandwidth must equal num_spectrum_resources * frequency_slot_bandwidth (qrmsa.pyx:294-295)"
        self.num_spectrum_resources = tb.n_slots
        self.k_paths = tb.k_paths
        self.modulations_to_consider = tb.mods_to_consider
        self.max_modulation_idx = tb.n_mods - 1
        self.bit_rates = tuple(bit_rates)
        self.frequency_start = frequency_start
        self.frequency_slot_bandwidth = frequency_slot_bandwidth
        self.frequency_end = frequency_end
        self.launch_power_dbm = launch_power_dbm
        self.launch_power = tb.launch_power_w
        self.margin = margin
        self.channel_width = channel_width
        self.action_space = Discrete(tb.n_actions)                                     # qrmsa.pyx:319-321
        self.observation_space = Box(-5, 5, (1 + 2 + tb.k_paths + tb.k_paths * tb.mods_to_consider * 12,))  # :323-335
        self.reject_action = self.action_space.n - 1
        if self.topology is not None:
            self.k_shortest_paths = self.topology.graph["ksp"]
            self.modulations = self.topology.graph.get("modulations", [])
            self._nodes = list(self.topology.graph["node_indices"])
        else:
            self.k_shortest_paths = None
            self.modulations = None
            self._nodes = list(tb.node_names) or [str(i) for i in range(tb.n_nodes)]


# ========================================================================================================
class QRMSAEnv(_Common):
    """Single environment, drop-in for the reference `QRMSAEnv` (constructor: qrmsa.pyx:206-237)."""

    def __init__(self, topology, num_spectrum_resources: int = 320, episode_length: int = 1000, load: float = 10.0,
                 mean_service_holding_time: float = 10800.0, bit_rate_selection: str = "continuous",
                 bit_rates: Sequence = (10, 40, 100), bit_rate_probabilities=None, node_request_probabilities=None,
                 bit_rate_lower_bound: float = 25.0, bit_rate_higher_bound: float = 100.0,
                 launch_power_dbm: float = 0.0, bandwidth: float = 4e12, frequency_start: float = (3e8 / 1565e-9),
                 frequency_slot_bandwidth: float = 12.5e9, margin: float = 0.0, measure_disruptions: bool = False,
                 seed=None, allow_rejection: bool = True, reset: bool = True, channel_width: float = 12.5,
                 k_paths: int = 5, file_name: str = "", blocks_to_consider: int = 1, modulations_to_consider: int = 6,
                 defragmentation: bool = False, n_defrag_services: int = 0, gen_observation: bool = True,
                 bands=None, device: int = 0):
        _check_kwargs(measure_disruptions, defragmentation, bands, gen_observation, bit_rate_selection)
        if seed is not None and not isinstance(seed, (int, np.integer)):
            raise ValueError("Seed must be an integer.")                       # qrmsa.pyx:342
        self.bit_rate_selection = bit_rate_selection
        randint_rates = None
        if bit_rate_selection == "continuous":
            bit_rates, randint_rates = _continuous_rates(bit_rate_lower_bound, bit_rate_higher_bound)
        self._setup(topology, num_spectrum_resources, bit_rates, launch_power_dbm, margin, frequency_start,
                    frequency_slot_bandwidth, channel_width, k_paths, modulations_to_consider, bandwidth)
        self.episode_length = int(episode_length)
        self.load = float(load)
        self.mean_service_holding_time = float(mean_service_holding_time)
        self.allow_rejection = allow_rejection
        self.gen_observation = gen_observation
        self.input_seed = int(seed) % (2 ** 31) if seed is not None else int.from_bytes(os.urandom(4), "little") >> 1
        tb = self.tables
        self._gen = TraceGenerator(1, tb.n_nodes, tb.n_rates, self.load, self.mean_service_holding_time,
                                   base_seed=self.input_seed, node_request_probabilities=node_request_probabilities,
                                   bit_rate_probabilities=bit_rate_probabilities, n_threads=1, randint_rates=randint_rates)
        # the device holds the next `_horizon` requests of the stream: one episode plus head-room for episodes that are
        # continued on the live network by reset(options={"only_episode_counters": True})
        self._horizon = int(min(16384, max(4 * self.episode_length, self.episode_length + 1, 2)))
        self._eng = Engine(tb, 1, self._horizon, device=device)
        self._eng.enable_gsnr_log(True)
        self.measure_disruptions, self.defragmentation = bool(measure_disruptions), bool(defragmentation)
        self.n_defrag_services = int(n_defrag_services)
        if self.measure_disruptions or self.defragmentation:
            self._eng.set_features(self.measure_disruptions, self.defragmentation, self.n_defrag_services)
        self.disrupted_services = self.episode_disrupted_services = 0
        self.episode_defrag_cicles = self.episode_service_realocations = 0
        self._running = []           # release keys of the running services (active_services column of the CSV)
        self.file_stats = None
        if file_name != "":          # qrmsa.pyx:387-406
            name = str(self.topology.graph["name"]) if self.topology is not None else "topology"
            self.final_file_name = "_".join([file_name, name, str(self.launch_power_dbm), str(self.load), str(seed) + ".csv"])
            d = os.path.dirname(self.final_file_name)
            if d and not os.path.exists(d):
                os.makedirs(d, exist_ok=True)
            self.file_stats = open(self.final_file_name, "wt", encoding="UTF-8")
            self.file_stats.write("# Service stats file from simulator\n")
            self.file_stats.write("id,source,destination,bit_rate,path_k,path_length,modulation,min_osnr,osnr,ase,nli,"
                                  "disrupted_services,active_services\n")
        self._block = None
        self._cur = 0
        self.current_time = 0.0
        self.current_service: Optional[Service] = None
        # cumulative counters (qrmsa.pyx:152-159) and per-episode ones
        self.services_processed = self.services_accepted = 0
        self.episode_services_processed = self.episode_services_accepted = 0
        self.bit_rate_requested = self.bit_rate_provisioned = 0.0
        self.episode_bit_rate_requested = self.episode_bit_rate_provisioned = 0.0
        self.bl_resource = self.bl_osnr = self.bl_reject = 0
        self.episode_modulation_histogram = {}
        import torch

        self._dev = torch.device("cuda", device)
        self._t_action = torch.zeros(1, dtype=torch.int64, device=self._dev)
        self._t_reward = torch.zeros(1, dtype=torch.float32, device=self._dev)
        self._t_status = torch.zeros(1, dtype=torch.uint8, device=self._dev)
        self._t_gsnr = torch.zeros(1, dtype=torch.float64, device=self._dev)
        self._t_term = torch.zeros(1, dtype=torch.uint8, device=self._dev)
        self._t_obs = torch.zeros((1, self.observation_space.shape[0]), dtype=torch.float32, device=self._dev)
        self._t_mask = torch.zeros((1, self.action_space.n), dtype=torch.uint8, device=self._dev)
        if reset:
            self.reset()

    # ------------------------------------------------------------------ gym API
    def _observation(self):
        if not self.gen_observation:                                           # qrmsa.pyx:584-587
            obs = np.zeros((self.observation_space.shape[0],), dtype=np.float32)
            return obs, {"mask": np.zeros((self.action_space.n,), dtype=np.uint8)}
        import torch

        self._eng.observation(self._t_obs, self._t_mask)
        torch.cuda.current_stream().synchronize()
        if self.tables.mods_to_consider < self.tables.n_mods:     # get_max_modulation_index, qrmsa.pyx:543-581, :680
            self.max_modulation_idx = int(self._eng.max_modulation_idx()[0])
        return self._t_obs[0].cpu().numpy(), {"mask": self._t_mask[0].cpu().numpy()}

    def _service_from_block(self, i: int) -> Service:
        b = self._block
        src, dst = int(b[0][i]), int(b[1][i])
        # service_id = episode_services_processed when the request is drawn (qrmsa.pyx:1092): the index within the episode
        return Service(service_id=self.episode_services_processed, source=self._nodes[src], source_id=src, destination=self._nodes[dst],
                       destination_id=dst, arrival_time=b[3][i], holding_time=b[4][i],
                       bit_rate=self.bit_rates[int(b[2][i])])

    def _account_new_service(self):
        svc = self.current_service
        self.current_time = svc.arrival_time
        now = np.float32(svc.arrival_time)            # releases due at the new request's arrival (qrmsa.pyx:1113-1122)
        self._running = [k for k in self._running if not (k <= now)]
        self.services_processed += 1
        self.episode_services_processed += 1
        self.bit_rate_requested += svc.bit_rate
        self.episode_bit_rate_requested += svc.bit_rate

    def reset(self, seed=None, options=None):
        """qrmsa.pyx:427-504.  Wipes the network, keeps the clock and the request stream running."""
        self.episode_services_processed = self.episode_services_accepted = 0
        self.episode_bit_rate_requested = self.episode_bit_rate_provisioned = 0.0
        self.bl_resource = self.bl_osnr = self.bl_reject = 0
        self.episode_modulation_histogram = {int(se): 0 for se in self.tables.mod_se}
        self.episode_disrupted_services = 0
        self.episode_defrag_cicles = self.episode_service_realocations = 0
        if options is not None and options.get("only_episode_counters"):
            if self.measure_disruptions or self.defragmentation:
                _unsupported("reset(options={'only_episode_counters': True}) together with measure_disruptions / defragmentation")
            # qrmsa.pyx:433, :461-464: counters restart, the network and the current request stay -- and because the
            # reference empties its release heap here, the services running now are never released
            if self._block is None:
                raise RuntimeError("reset(options={'only_episode_counters': True}) before the first full reset")
            self._eng.cancel_pending_releases()
            self._running = []
            obs, _ = self._observation()
            return obs, {}
        L = self._horizon
        if self._block is None:
            leftover = [np.zeros(0, a) for a in (np.uint8, np.uint8, np.uint8, np.float32, np.float32)]
        else:  # requests generated but never made current keep their place in the stream
            leftover = [a[self._cur + 1:] for a in self._block]
        need = L - len(leftover[0])
        fresh = [a[:, 0] for a in self._gen.next(max(need, 0))]
        self._block = [np.concatenate([lo, fr])[:L] for lo, fr in zip(leftover, fresh)]
        self._eng.reset()
        self._eng.load_trace_host(*[np.ascontiguousarray(a[:, None]) for a in self._block])
        self._cur = 0
        self.max_modulation_idx = self.tables.n_mods - 1                      # :437
        self._running = []
        self.bit_rate_requested = self.bit_rate_provisioned = 0.0            # :466-467 (full reset)
        self.disrupted_services = 0                                          # :468
        self.current_service = self._service_from_block(0)
        self._account_new_service()
        obs, mask = self._observation()
        return obs, mask.copy()

    def step(self, action: int):
        """qrmsa.pyx:838-1065."""
        import torch

        tb = self.tables
        svc = self.current_service
        svc.blocked_due_to_resources = svc.blocked_due_to_osnr = False
        disrupted_now = 0
        self._t_action[0] = int(action)
        self._eng.step_action(self._t_action, self._t_reward, self._t_status, self._t_gsnr, self._t_term)
        torch.cuda.current_stream().synchronize()
        status = int(self._t_status[0])
        reward = float(self._t_reward[0])
        gsnr = float(self._t_gsnr[0])
        if status == _lib.STEP_IDLE:
            raise RuntimeError("step() called after the episode terminated; call reset()")
        route = modulation_idx = initial_slot = -1
        osnr_req = 0.0
        if action != self.reject_action:
            initial_slot, modulation_idx, route = self.encoded_decimal_to_array(int(action))[::-1]
            osnr_req = float(tb.mod_min_osnr[modulation_idx]) + self.margin
        if status == _lib.STEP_NOT_FREE:                                      # :886-897: request not consumed
            svc.blocked_due_to_resources, svc.accepted = True, False
            obs, mask = self._observation()
            info = {"blocked_due_to_resources": 1, "blocked_due_to_osnr": 0, "rejected": 1}
            info.update(mask)
            return obs, reward, False, False, info
        if status == _lib.STEP_LOW_GSNR:                                      # :925-929
            raise ValueError(f"Osnr {gsnr} is not enough for service {svc.service_id} with modulation index "
                             f"{modulation_idx}, and osnr_req {osnr_req}.")
        if status == _lib.STEP_ACCEPTED:
            n = int(tb.slots_needed[int(self._block[2][self._cur]) * tb.n_mods + modulation_idx])
            svc.accepted = True
            svc.OSNR = gsnr
            ase, nli = self._eng.ase_nli_host(self._cur, 1)                    # osnr.pyx:133-140, stored at :930-932
            svc.ASE, svc.NLI = float(ase[0, 0]), float(nli[0, 0])
            self._running.append(np.float32(np.float32(svc.arrival_time) + np.float32(svc.holding_time)))   # :1327-1330
            if self.measure_disruptions:                                       # :937-952
                disrupted_now = int(self._eng.step_disrupted()[0])
                self.disrupted_services += disrupted_now
                self.episode_disrupted_services += disrupted_now
            svc.initial_slot, svc.number_slots = initial_slot, n
            svc.center_frequency = (self.frequency_start + (self.frequency_slot_bandwidth * initial_slot)
                                    + (self.frequency_slot_bandwidth * (n / 2.0)))
            svc.bandwidth = self.frequency_slot_bandwidth * n
            svc.launch_power = self.launch_power
            if self.k_shortest_paths is not None:
                svc.path = self.k_shortest_paths[svc.source, svc.destination][route]
                svc.current_modulation = self.modulations[modulation_idx]
            self.services_accepted += 1
            self.episode_services_accepted += 1
            self.bit_rate_provisioned += svc.bit_rate
            self.episode_bit_rate_provisioned = float(int(self.episode_bit_rate_provisioned + svc.bit_rate))  # :1319
            self.episode_modulation_histogram[int(tb.mod_se[modulation_idx])] += 1
        else:
            svc.accepted = False
            self.bl_reject += 1
        if self.file_stats is not None:                                        # qrmsa.pyx:967-990
            line = "{},{},{},{},".format(svc.service_id, svc.source_id, svc.destination_id, svc.bit_rate)
            if svc.accepted:
                if self.k_shortest_paths is not None:
                    pk, plen = svc.path.k, svc.path.length
                    se, mo = svc.current_modulation.spectral_efficiency, svc.current_modulation.minimum_osnr
                else:
                    pk = route
                    plen = float(tb.path_length_km[tb.path_index(svc.source_id, int(svc.destination_id), route)])
                    se, mo = int(tb.mod_se[modulation_idx]), float(tb.mod_min_osnr[modulation_idx])
                line += "{},{},{},{},{},{},{},{},{}".format(pk, plen, se, mo, svc.OSNR, svc.ASE, svc.NLI, disrupted_now, len(self._running))
            else:
                line += "-1,-1,-1,-1,-1,-1,-1,-1,-1"
            self.file_stats.write(line + "\n")
            self.file_stats.flush()
        info = {
            "episode_services_accepted": self.episode_services_accepted,
            "service_blocking_rate": 0.0, "episode_service_blocking_rate": 0.0,
            "bit_rate_blocking_rate": 0.0, "episode_bit_rate_blocking_rate": 0.0,
            "disrupted_services": 0.0, "episode_disrupted_services": 0.0,
            "osnr": gsnr if status == _lib.STEP_ACCEPTED else 0.0, "osnr_req": osnr_req,
            "chosen_path_index": route, "chosen_slot": initial_slot,
            # assembled before the next request is drawn (qrmsa.pyx:996-1052): the release phase of THIS step shows up
            # in the next step's info
            "episode_defrag_cicles": self.episode_defrag_cicles,
            "episode_service_realocations": self.episode_service_realocations,
        }
        if self.disrupted_services > 0 and self.services_accepted > 0:            # :1035-1036
            info["disrupted_services"] = float(self.disrupted_services) / self.services_accepted
        if self.episode_disrupted_services > 0 and self.episode_services_accepted > 0:
            # :1038-1041 -- int / int in C before the float(): the quotient is truncated
            info["episode_disrupted_services"] = float(self.episode_disrupted_services // self.episode_services_accepted)
        if self.defragmentation:
            c = self._eng.counters()[0]
            self.episode_defrag_cicles, self.episode_service_realocations = int(c[26]), int(c[27])
        if self.services_processed > 0:
            info["service_blocking_rate"] = float(self.services_processed - self.services_accepted) / self.services_processed
        if self.episode_services_processed > 0:
            info["episode_service_blocking_rate"] = (float(self.episode_services_processed - self.episode_services_accepted)
                                                     / float(self.episode_services_processed))
        if self.bit_rate_requested > 0:
            info["bit_rate_blocking_rate"] = float(self.bit_rate_requested - self.bit_rate_provisioned) / self.bit_rate_requested
        if self.episode_bit_rate_requested > 0:
            info["episode_bit_rate_blocking_rate"] = (float(self.episode_bit_rate_requested - self.episode_bit_rate_provisioned)
                                                      / self.episode_bit_rate_requested)
        for se in self.tables.mod_se:
            info["modulation_{}".format(str(float(se)))] = self.episode_modulation_histogram.get(int(se), 0)
        # next request (qrmsa.pyx:1052-1054)
        self._cur += 1
        if self._cur >= len(self._block[0]) - 1:
            raise RuntimeError(f"the {self._horizon} requests loaded on the device are used up (episodes continued with "
                               "reset(options={'only_episode_counters': True})); a full reset() attaches the next ones")
        self.current_service = self._service_from_block(self._cur)
        self._account_new_service()
        terminated = self.episode_services_processed == self.episode_length
        if terminated:
            info["blocked_due_to_resources"] = self.bl_resource
            info["blocked_due_to_osnr"] = self.bl_osnr
            info["rejected"] = self.bl_reject
        obs, mask = self._observation()
        info.update(mask)
        return obs, reward, terminated, False, info

    # ------------------------------------------------------------------ heuristics' call surface
    def encoded_decimal_to_array(self, decimal: int, max_values=None):
        """qrmsa.pyx:801-834 -> [route, modulation index, initial slot]."""
        if max_values is None:
            max_values = [self.k_paths, self.modulations_to_consider, self.num_spectrum_resources]
        arr = []
        for mv in reversed(max_values):
            arr.insert(0, decimal % mv)
            decimal //= mv
        if self.max_modulation_idx > 1:
            allowed = list(range(self.max_modulation_idx, self.max_modulation_idx - self.modulations_to_consider, -1))
        else:
            allowed = list(reversed(range(0, self.modulations_to_consider)))
        arr[1] = allowed[arr[1]]
        return arr

    def get_number_slots(self, service, modulation) -> int:
        """qrmsa.pyx:1198-1205 (bands=None)."""
        return int(math.ceil(service.bit_rate / (modulation.spectral_efficiency * self.channel_width)))

    def _path_link_indices(self, path):
        nl = path.node_list
        return [int(self.topology[nl[i]][nl[i + 1]]["index"]) for i in range(len(nl) - 1)]

    def available_slots_matrix(self) -> np.ndarray:
        """topology.graph['available_slots'] equivalent: int32 [E][S], 1 = free."""
        return self._eng.export_slots(0)

    def get_available_slots(self, path) -> np.ndarray:
        """qrmsa.pyx:1482-1512."""
        m = self.available_slots_matrix()
        out = m[self._path_link_indices(path)[0]].copy()
        for l in self._path_link_indices(path)[1:]:
            out *= m[l]
        return out

    def _get_spectrum_slots(self, path_idx: int):
        """qrmsa.pyx:1533-1542."""
        m = self.available_slots_matrix()
        svc = self.current_service
        return [m[l] for l in self._path_link_indices(self.k_shortest_paths[svc.source, svc.destination][path_idx])]

    def _get_candidates(self, available_slots, num_slots_required: int, total_slots: int):
        """qrmsa.pyx:515-541: valid start slots under the guard-band rule."""
        av = np.asarray(available_slots).astype(np.int8)
        edges = np.flatnonzero(np.diff(np.concatenate([[0], av, [0]])))
        out = []
        for start, end in zip(edges[::2], edges[1::2]):
            length = end - start
            if start + length == total_slots:
                if length >= num_slots_required:
                    out.extend(range(start, start + length - num_slots_required + 1))
            elif length >= num_slots_required + 1:
                out.extend(range(start, start + length - (num_slots_required + 1) + 1))
        return [int(x) for x in out]

    def is_path_free(self, path, initial_slot: int, number_slots: int) -> bool:
        """qrmsa.pyx:1248-1264."""
        end = initial_slot + number_slots
        if end > self.num_spectrum_resources:
            return False
        if end < self.num_spectrum_resources:
            end += 1
        m = self.available_slots_matrix()
        return all(not np.any(m[l, initial_slot:end] == 0) for l in self._path_link_indices(path))

    def running_services_on_link(self, link_index: int) -> np.ndarray:
        """Channels on a link as rows (initial_slot, number_slots, modulation index)."""
        return self._eng.export_link_list(0, link_index)

    def calculate_osnr(self, service):
        """core.osnr.calculate_osnr(env, service) (osnr.pyx:21-142) for the candidate written on `service` (path,
        initial_slot, number_slots): (GSNR, ASE-only, NLI-only) in dB."""
        paths = self.k_shortest_paths[service.source, service.destination]
        p = next(i for i, pth in enumerate(paths) if pth is service.path)
        return self._eng.probe_qot(0, service.source_id, int(service.destination_id), p, service.initial_slot,
                                   service.number_slots)

    def close(self):
        self._eng.close()
        self._gen.close()
        if self.file_stats is not None:
            self.file_stats.close()
            self.file_stats = None


def calculate_osnr(env, service):
    """Module-level form used by the reference heuristics: `calculate_osnr(env, service)` (osnr.pyx:21)."""
    while not isinstance(env, QRMSAEnv) and hasattr(env, "env"):
        env = env.env
    return env.calculate_osnr(service)


# ========================================================================================================
class BatchedQRMSAEnv(_Common):
    """n_envs environments on one GPU, stepped together.  `load` may be a scalar or one value per env (load sweeps:
    set n_groups for per-load counters).

    request_source="replay" (default): env i replays random.Random(base_seed + i), generated on the host draw for
    draw as the reference does (qrmsa.pyx:1079-1099) and uploaded each episode.
    request_source="device": the same traffic model drawn on the GPU from Philox streams keyed by
    (seed, env_offset + i) -- no host generation or upload; not the reference's streams."""

    def __init__(self, topology, n_envs: int, num_spectrum_resources: int = 320, episode_length: int = 1000,
                 load=10.0, mean_service_holding_time: float = 10800.0, bit_rate_selection: str = "discrete",
                 bit_rates: Sequence = (10, 40, 100), bit_rate_probabilities=None, node_request_probabilities=None,
                 launch_power_dbm: float = 0.0, bandwidth: float = 4e12, frequency_start: float = (3e8 / 1565e-9),
                 frequency_slot_bandwidth: float = 12.5e9, margin: float = 0.0, measure_disruptions: bool = False,
                 seed: int = 50, allow_rejection: bool = True, reset: bool = True, channel_width: float = 12.5,
                 k_paths: int = 5, modulations_to_consider: int = 6, defragmentation: bool = False,
                 n_defrag_services: int = 0,
                 gen_observation: bool = False, bands=None, device: int = 0, n_groups: int = 1, n_threads: int = 0,
                 request_source: str = "replay", env_offset: int = 0):
        _check_kwargs(measure_disruptions, defragmentation, bands, gen_observation, bit_rate_selection)
        if bit_rate_selection != "discrete":
            _unsupported("bit_rate_selection='continuous' on BatchedQRMSAEnv (QRMSAEnv supports it)")
        self._setup(topology, num_spectrum_resources, bit_rates, launch_power_dbm, margin, frequency_start,
                    frequency_slot_bandwidth, channel_width, k_paths, modulations_to_consider, bandwidth)
        import torch

        if request_source not in ("replay", "device"):
            raise ValueError("request_source must be 'replay' or 'device'")
        self.request_source = request_source
        self.env_offset = int(env_offset)
        self._traffic = dict(load=load, mean_holding_time=mean_service_holding_time,
                             node_request_probabilities=node_request_probabilities,
                             bit_rate_probabilities=bit_rate_probabilities)
        self._episodes = 0
        self.n_envs = int(n_envs)
        self.episode_length = int(episode_length)
        self.base_seed = int(seed)
        tb = self.tables
        self._gen = TraceGenerator(self.n_envs, tb.n_nodes, tb.n_rates, load, mean_service_holding_time,
                                   base_seed=self.base_seed, node_request_probabilities=node_request_probabilities,
                                   bit_rate_probabilities=bit_rate_probabilities, n_threads=n_threads)
        self._eng = Engine(tb, self.n_envs, max(self.episode_length, 2), device=device)
        if n_groups > 1:
            self._eng.set_groups(n_groups)
        if measure_disruptions or defragmentation:
            self._eng.set_features(measure_disruptions, defragmentation, n_defrag_services)
        self._dev = torch.device("cuda", device)
        shape = (self.episode_length, self.n_envs)
        self._pinned = [torch.empty(shape, dtype=dt, pin_memory=True) for dt in
                        (torch.uint8, torch.uint8, torch.uint8, torch.float32, torch.float32)]
        self._trace = [p.numpy() for p in self._pinned]
        self._reward = torch.zeros(self.n_envs, dtype=torch.float32, device=self._dev)
        self._status = torch.zeros(self.n_envs, dtype=torch.uint8, device=self._dev)
        self._gsnr = torch.zeros(self.n_envs, dtype=torch.float64, device=self._dev)
        self._term = torch.zeros(self.n_envs, dtype=torch.uint8, device=self._dev)
        self._obs = torch.zeros((self.n_envs, self.observation_space.shape[0]), dtype=torch.float32, device=self._dev)
        self.gen_observation = bool(gen_observation)
        self._mask = (torch.zeros((self.n_envs, self.action_space.n), dtype=torch.uint8, device=self._dev)
                      if self.gen_observation else None)
        self.steps_done = 0
        if reset:
            self.reset()

    @property
    def engine(self) -> Engine:
        return self._eng

    def reset(self, seed=None, options=None):
        """Every env: network wiped, next `episode_length` requests of its stream attached (qrmsa.pyx:427-504)."""
        self._eng.reset()
        if self.request_source == "device":
            self._eng.generate_trace(self.episode_length, self._traffic["load"], seed=self.base_seed,
                                     restart=self._episodes == 0, env_offset=self.env_offset,
                                     mean_holding_time=self._traffic["mean_holding_time"],
                                     node_request_probabilities=self._traffic["node_request_probabilities"],
                                     bit_rate_probabilities=self._traffic["bit_rate_probabilities"])
            self._trace_valid = False
        else:
            self._gen.next(self.episode_length, out=self._trace)
            self._eng.load_trace_host(*self._trace)
            self._trace_valid = True
        self._episodes += 1
        self.steps_done = 0
        self._rl_path = False
        self._term.zero_()
        self._observe()
        return self._obs, {"mask": self._mask}

    def _observe(self):
        if self.gen_observation:
            self._eng.observation(self._obs, self._mask)

    def current_requests(self):
        """(src, dst, rate index, arrival, holding) host arrays of the episode, [episode_length, n_envs]."""
        if not self._trace_valid:   # device-generated: fetched on demand
            for dst_arr, a in zip(self._trace, self._eng.trace_host()):
                dst_arr[...] = a
            self._trace_valid = True
        return self._trace

    def step(self, actions):
        """actions: int64 CUDA tensor [n_envs] -> (obs, reward, terminated, truncated, info) of device tensors.
        An env whose action is refused (status NOT_FREE / LOW_GSNR) keeps its request, as in the reference
        (qrmsa.pyx:886-897), so after such calls the envs are at different requests: `steps_done` counts CALLS, the
        per-env progress is `engine.env_state()[:, 0]` and the per-env end of episode is the returned `terminated`."""
        self._eng.step_action(actions, self._reward, self._status, self._gsnr, self._term)
        self.steps_done += 1
        self._rl_path = True
        self._observe()
        info = {"status": self._status, "osnr": self._gsnr, "mask": self._mask}
        return self._obs, self._reward, self._term.bool(), self._term.bool() & False, info

    def step_first_fit(self, n_steps: int = 1):
        """n_steps of `heuristic_shortest_available_path_first_fit_best_modulation` + `env.step` per env."""
        n_steps = min(int(n_steps), self.episode_length - 1 - self.steps_done)
        self._eng.step_first_fit(n_steps)
        self.steps_done += n_steps
        if n_steps:
            self._observe()
        return n_steps

    def step_heuristic(self, policy: str, n_steps: int = 1):
        """Fused device policy + step: "first_fit" (heuristics.py:923) or "load_balancing" (heuristics.py:547)."""
        n_steps = min(int(n_steps), self.episode_length - 1 - self.steps_done)
        self._eng.step_heuristic(policy, n_steps)
        self.steps_done += n_steps
        if n_steps:
            self._observe()
        return n_steps

    @property
    def terminated(self) -> bool:
        """Every env has decided its last request.  After step(actions) calls the envs may have progressed unevenly
        (refused actions do not consume a request): the device's per-env flags decide then."""
        if getattr(self, "_rl_path", False):
            return bool(self._term.all())
        return self.steps_done >= self.episode_length - 1

    def action_masks(self):
        """Last action mask, uint8 CUDA tensor [n_envs, k*M*S+1] (wrappers/qrmsa_gym.py:74-75); None when
        gen_observation=False (the reference returns zeros there)."""
        return self._mask

    def actions(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """Decided action indices [count, n_envs] (reject = k*M*S) and near-threshold flags."""
        count = self.steps_done - first if count is None else count
        w = self._eng.actions_host(first, count)
        return (w & _lib.ACTION_MASK).astype(np.int64), (w.view(np.uint32) & _lib.FLAG_NEAR_THRESHOLD) != 0

    def counters(self, group: Optional[int] = None) -> dict:
        return self._eng.counters_dict(group)

    def episode_info(self, group: Optional[int] = None) -> dict:
        """Blocking statistics in the reference's `info` vocabulary (qrmsa.pyx:996-1050), aggregated."""
        c = self.counters(group)
        dec = max(c["decided"], 1)
        req = max(c["rate_requested_milli"], 1)
        return {
            "episode_services_processed": c["decided"], "episode_services_accepted": c["accepted"],
            "episode_service_blocking_rate": (c["decided"] - c["accepted"]) / dec,
            "episode_bit_rate_blocking_rate": (c["rate_requested_milli"] - c["rate_provisioned_milli"]) / req,
            "rejected": c["rejected"], "near_threshold_decisions": c["near_threshold"],
            "disrupted_services": c["disrupted_services"] / max(c["accepted"], 1),
            "episode_defrag_cicles": c["defrag_cycles"], "episode_service_realocations": c["service_reallocations"],
            **{f"modulation_{float(se)}": int(c["mod_hist"][i]) for i, se in enumerate(self.tables.mod_se)},
        }

    def available_slots(self, env: int) -> np.ndarray:
        return self._eng.export_slots(env)

    def bitmaps(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.n_envs - first if count is None else count
        return unpack_bitmaps(self._eng.export_bitmaps(first, count), self.tables.n_slots)

    def close(self):
        self._eng.close()
        self._gen.close()
