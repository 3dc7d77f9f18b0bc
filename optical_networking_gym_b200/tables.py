"""Dense static tables exported from a reference `topology` graph.

The reference keeps its network as a networkx graph of Python objects
(reference optical_networking_gym/topology.pyx:244-369): `graph["ksp"][(n1, n2)]`
is a list of k `Path` objects with `.links` (tuple of `Link` with `.node1/.node2/
.spans`), edges carry `"index"`/`"link"`, `graph["modulations"]` the modulation
tuple and `graph["node_indices"]` the node order.  The device kernels need the
same information as flat arrays.  The k-shortest-path ORDER depends on networkx
tie-breaking (topology.pyx:100-104), so the tables are always exported from the
live object (duck-typed: nothing is imported from the reference), never
recomputed; exported tables can be saved to / loaded from `.npz` so GPU runs do
not need the topology files.

Slots needed per (bit-rate, modulation) follow `QRMSAEnv.get_number_slots`
(reference envs/qrmsa.pyx:1198-1205): ceil(bit_rate / (SE * channel_width)).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Sequence

import numpy as np


@dataclasses.dataclass
class StaticTables:
    name: str
    n_nodes: int
    n_links: int
    k_paths: int
    n_mods: int
    mods_to_consider: int
    n_rates: int
    n_slots: int
    max_hops: int
    path_hops: np.ndarray        # u8  [n_nodes*n_nodes*k]
    path_links: np.ndarray       # u8  [n_nodes*n_nodes*k*max_hops]
    path_length_km: np.ndarray   # f64 [n_nodes*n_nodes*k]
    link_n_spans: np.ndarray     # i32 [E]
    link_span_len_m: np.ndarray  # f64 [E]
    link_alpha: np.ndarray       # f64 [E]  attenuation_normalized (1/m)
    link_nf: np.ndarray          # f64 [E]  noise figure, linear
    mod_se: np.ndarray           # i32 [M]
    mod_min_osnr: np.ndarray     # f64 [M]
    bit_rates: np.ndarray        # f64 [R]
    slots_needed: np.ndarray     # u8  [R*M]
    frequency_start: float
    slot_bandwidth_hz: float
    launch_power_w: float
    margin_db: float
    node_names: tuple = ()

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_topology(
        cls,
        topology,
        num_spectrum_resources: int = 320,
        bit_rates: Sequence[float] = (10, 40, 100),
        launch_power_dbm: float = 0.0,
        margin: float = 0.0,
        frequency_start: float = 3e8 / 1565e-9,
        frequency_slot_bandwidth: float = 12.5e9,
        channel_width: float = 12.5,
        k_paths: int = 5,
        modulations_to_consider: int = 6,
    ) -> "StaticTables":
        g = topology.graph
        nodes = list(g["node_indices"])
        n_nodes = len(nodes)
        ksp = g["ksp"]
        mods = list(g.get("modulations", []))
        n_mods = len(mods)
        n_links = topology.number_of_edges()
        if n_links > 255:
            raise ValueError("link indices are stored as uint8 (n_links <= 255)")

        link_n_spans = np.zeros(n_links, np.int32)
        link_span_len_m = np.zeros(n_links, np.float64)
        link_alpha = np.zeros(n_links, np.float64)
        link_nf = np.zeros(n_links, np.float64)
        seen = set()
        for u, v in topology.edges():
            idx = int(topology[u][v]["index"])
            spans = topology[u][v]["link"].spans
            s0 = spans[0]
            for sp in spans:  # the factorised GN evaluation relies on identical spans per link (topology.pyx:288-299)
                if (sp.length != s0.length or sp.attenuation_normalized != s0.attenuation_normalized
                        or sp.noise_figure_normalized != s0.noise_figure_normalized):
                    raise ValueError("spans of one link must be identical")
            link_n_spans[idx] = len(spans)
            link_span_len_m[idx] = s0.length * 1e3
            link_alpha[idx] = s0.attenuation_normalized
            link_nf[idx] = s0.noise_figure_normalized
            seen.add(idx)
        if seen != set(range(n_links)):
            raise ValueError("edge 'index' attributes must be 0..E-1")

        max_hops = 1
        for paths in ksp.values():
            for p in paths[:k_paths]:
                max_hops = max(max_hops, len(p.links))
        n_paths = n_nodes * n_nodes * k_paths
        path_hops = np.zeros(n_paths, np.uint8)
        path_links = np.zeros((n_paths, max_hops), np.uint8)
        path_length = np.zeros(n_paths, np.float64)
        for (a, b), paths in ksp.items():
            ia, ib = nodes.index(a), nodes.index(b)
            for p_i, p in enumerate(paths[:k_paths]):
                pi = (ia * n_nodes + ib) * k_paths + p_i
                path_hops[pi] = len(p.links)
                path_length[pi] = p.length
                for h, ln in enumerate(p.links):
                    path_links[pi, h] = int(topology[ln.node1][ln.node2]["index"])

        rates = np.asarray(bit_rates, np.float64)
        need = np.zeros((len(rates), n_mods), np.int64)
        for r, rate in enumerate(rates):
            for m, mod in enumerate(mods):
                # bit_rate is a C float in the reference Service (qrmsa.pyx:37)
                need[r, m] = int(math.ceil(float(np.float32(rate)) / (mod.spectral_efficiency * channel_width)))
        if need.max() > 255 or need.min() < 1:
            raise ValueError("slots needed must be in 1..255")

        return cls(
            name=str(g.get("name", "topology")),
            n_nodes=n_nodes, n_links=n_links, k_paths=int(k_paths), n_mods=n_mods,
            mods_to_consider=min(int(modulations_to_consider), n_mods),
            n_rates=len(rates), n_slots=int(num_spectrum_resources), max_hops=int(max_hops),
            path_hops=path_hops, path_links=path_links.reshape(-1), path_length_km=path_length,
            link_n_spans=link_n_spans, link_span_len_m=link_span_len_m, link_alpha=link_alpha, link_nf=link_nf,
            mod_se=np.array([m.spectral_efficiency for m in mods], np.int32),
            mod_min_osnr=np.array([m.minimum_osnr for m in mods], np.float64),
            bit_rates=rates, slots_needed=need.astype(np.uint8).reshape(-1),
            frequency_start=float(frequency_start), slot_bandwidth_hz=float(frequency_slot_bandwidth),
            launch_power_w=float(10 ** ((launch_power_dbm - 30) / 10)),  # qrmsa.pyx:288
            margin_db=float(margin), node_names=tuple(str(n) for n in nodes),
        )

    # ------------------------------------------------------------------ persistence
    _ARRAYS = ("path_hops", "path_links", "path_length_km", "link_n_spans", "link_span_len_m", "link_alpha",
               "link_nf", "mod_se", "mod_min_osnr", "bit_rates", "slots_needed")
    _SCALARS = ("n_nodes", "n_links", "k_paths", "n_mods", "mods_to_consider", "n_rates", "n_slots", "max_hops")
    _FLOATS = ("frequency_start", "slot_bandwidth_hz", "launch_power_w", "margin_db")

    def save(self, path: str) -> None:
        d = {k: getattr(self, k) for k in self._ARRAYS}
        d.update({k: np.int64(getattr(self, k)) for k in self._SCALARS})
        d.update({k: np.float64(getattr(self, k)) for k in self._FLOATS})
        d["name"] = np.array(self.name)
        d["node_names"] = np.array(list(self.node_names))
        np.savez_compressed(path, **d)

    @classmethod
    def load(cls, path: str) -> "StaticTables":
        z = np.load(path, allow_pickle=False)
        kw = {k: np.ascontiguousarray(z[k]) for k in cls._ARRAYS}
        kw.update({k: int(z[k]) for k in cls._SCALARS})
        kw.update({k: float(z[k]) for k in cls._FLOATS})
        return cls(name=str(z["name"]), node_names=tuple(str(s) for s in z["node_names"]), **kw)

    def replace(self, **kw) -> "StaticTables":
        return dataclasses.replace(self, **kw)

    # ------------------------------------------------------------------ helpers
    @property
    def n_actions(self) -> int:
        """k * M * S + 1; the last action is 'reject' (qrmsa.pyx:319-321)."""
        return self.k_paths * self.mods_to_consider * self.n_slots + 1

    def path_index(self, src: int, dst: int, p: int = 0) -> int:
        return (src * self.n_nodes + dst) * self.k_paths + p

    def links_of(self, src: int, dst: int, p: int) -> np.ndarray:
        pi = self.path_index(src, dst, p)
        return self.path_links.reshape(-1, self.max_hops)[pi, : self.path_hops[pi]]
