"""Search space of the descriptor supernet and its per-layer / per-candidate latency table
(hardnetNAS/supernet_functions/lookup_table_builder.py:18-45 search space, :78-110 constructor arguments, :121-158 latency
measurement, :160-188 the `lookup_table.txt` text format).

The table is search-time tooling: the supernet's latency loss sums `lookup_table_latency[layer][op]` over the sampled ops. The
reference times every candidate as a stock torch module at batch 1000 with wall-clock time. Here the same table can be measured
two ways, both with CUDA events:
  engine="torch"  the reference's procedure (the candidate's torch module on the current CUDA device), and
  engine="b200"   what the candidate costs on THIS library's kernels: the eval forward of a `SampledDescriptorNet` with the
                  candidate at that layer and `skip` everywhere else, minus the all-`skip` net (differential, because the engine
                  fuses across layer boundaries: a candidate's cost depends on what it is fused with).
Nothing in the descriptor hot path reads the table."""
from collections import OrderedDict

from .fbnet_builder import PRIMITIVES

CANDIDATE_BLOCKS = ["skip", "ir_k3_e1", "ir_k3_e3", "ir_k3_s4", "ir_k5_e1", "ir_k5_e3", "ir_k5_s4", "ir_k3_e1_se",
                    "ir_k3_e3_se", "ir_k3_s4_se", "ir_k5_e1_se", "ir_k5_e3_se", "ir_k5_s4_se", "ir_k3_s2", "ir_k5_s2",
                    "ir_k3_s2_se", "ir_k5_s2_se"]

SEARCH_SPACE2 = OrderedDict([
    ("input_shape", [(32, 32, 32), (32, 16, 16), (32, 16, 16), (64, 8, 8), (64, 8, 8), (128, 4, 4)]),
    ("channel_size", [32, 32, 64, 64, 128, 128]),
    ("strides", [2, 1, 2, 1, 2, 1]),
])


class LookUpTable:
    """Per-layer constructor arguments (C_in, C_out, -999, stride) and the candidate op constructors."""

    def __init__(self, candidate_blocks=CANDIDATE_BLOCKS, search_space=SEARCH_SPACE2, calulate_latency=False,
                 path_to_file=None, cnt_of_runs=50, engine="b200"):
        """`calulate_latency` (the reference's spelling) measures the table and, with `path_to_file`, writes it; otherwise a
        given `path_to_file` is read. Without either the table stays None (the deployed nets never need it)."""
        self.cnt_layers = len(search_space["input_shape"])
        self.lookup_table_operations = {name: PRIMITIVES[name] for name in candidate_blocks}
        self.layers_parameters = [(search_space["input_shape"][i][0], search_space["channel_size"][i], -999,
                                   search_space["strides"][i]) for i in range(self.cnt_layers)]
        self.layers_input_shapes = search_space["input_shape"]
        self.lookup_table_latency = None
        if calulate_latency:
            self._create_from_operations(cnt_of_runs, write_to_file=path_to_file, engine=engine)
        elif path_to_file is not None:
            self._create_from_file(path_to_file)

    # ---- measurement (lookup_table_builder.py:113-158) ---------------------------------------------------------
    def _create_from_operations(self, cnt_of_runs, write_to_file=None, engine="b200"):
        self.lookup_table_latency = self._calculate_latency(self.lookup_table_operations, self.layers_parameters,
                                                            self.layers_input_shapes, cnt_of_runs, engine)
        if write_to_file is not None:
            self._write_lookup_table_to_file(write_to_file)

    @staticmethod
    def _time_ms(fn, cnt_of_runs):
        import torch
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(cnt_of_runs):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / cnt_of_runs

    def _calculate_latency(self, operations, layers_parameters, layers_input_shapes, cnt_of_runs, engine="b200"):
        """Milliseconds per batch of 1000 (the reference's unit: `total_time / cnt_of_runs * 1e3` of seconds)."""
        import torch
        LATENCY_BATCH_SIZE = 1000
        if not torch.cuda.is_available():
            raise RuntimeError("the latency table is measured on a CUDA device")
        table = [{} for _ in range(self.cnt_layers)]
        if engine == "torch":
            for layer_id in range(self.cnt_layers):
                x = torch.randn((LATENCY_BATCH_SIZE, *layers_input_shapes[layer_id]), device="cuda")
                for op_name, ctor in operations.items():
                    op = ctor(*layers_parameters[layer_id]).cuda().eval()
                    with torch.no_grad():
                        table[layer_id][op_name] = self._time_ms(lambda: op(x), cnt_of_runs)
            return table
        if engine != "b200":
            raise ValueError("engine must be 'b200' or 'torch'")
        from .descriptor_net import SampledDescriptorNet
        if self.cnt_layers != 6 or "skip" not in PRIMITIVES:
            raise ValueError("the engine measurement builds whole descriptor nets: it needs the six-layer SEARCH_SPACE2")
        x = torch.rand((LATENCY_BATCH_SIZE, 1, 32, 32), device="cuda")

        def net_ms(ops):
            net = SampledDescriptorNet(ops, chunk_patches=1024, head_rows=1024).cuda().eval()
            out = torch.empty((LATENCY_BATCH_SIZE, 128), device="cuda")
            return self._time_ms(lambda: net(x, out=out), cnt_of_runs)

        base = net_ms(["skip"] * 6)
        for layer_id in range(self.cnt_layers):
            for op_name in operations:
                if op_name == "skip":
                    table[layer_id][op_name] = 0.0
                    continue
                ops = ["skip"] * 6
                ops[layer_id] = op_name
                table[layer_id][op_name] = max(net_ms(ops) - base, 0.0)
        return table

    # ---- text format (lookup_table_builder.py:160-188): op names on the first line, one line of latencies per layer -------
    def _write_lookup_table_to_file(self, path_to_file):
        ops = list(self.lookup_table_operations)
        lines = [" ".join(ops)]
        for layer_id in range(self.cnt_layers):
            lines.append(" ".join(str(self.lookup_table_latency[layer_id][op]) for op in ops))
        with open(path_to_file, "w") as f:
            f.write("\n".join(lines))

    def _create_from_file(self, path_to_file):
        self.lookup_table_latency = self._read_lookup_table_from_file(path_to_file)

    def _read_lookup_table_from_file(self, path_to_file):
        lines = [line.strip("\n") for line in open(path_to_file)]
        ops_names = lines[0].split(" ")
        rows = [list(map(float, layer.split(" "))) for layer in lines[1:] if layer.strip()]
        return [{op: rows[i][k] for k, op in enumerate(ops_names)} for i in range(self.cnt_layers)]
