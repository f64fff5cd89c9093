"""CPU oracle for the raincast-gnn training hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy for the integer/graph work, plain CPU
PyTorch ops for the floating-point work) of the reference algorithm on the hot
path SURVEY.md section 8 names:

    graph construction  -> oracle/graph.py   (utils/data.py:261-284 + PyG collate)
    GINEConv / batching -> oracle/pyg.py     (torch_geometric, NOT vendored in the reference)
    DeepSets / ResGnn   -> oracle/model.py   (models/gnn.py:10-141)
    losses / links      -> oracle/losses.py  (models/loss.py, models/model_utils.py)

Who may import it: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline /
`--impl reference` legs.  Nothing under raincast_gnn_b200/ imports it; the product
path has no CPU fallback and raises when the CUDA library is missing.

Pinning.  The reference ships no tests, golden vectors or checkpoints
(SURVEY.md section 4), and its GINEConv arithmetic lives in `torch-geometric`
(unpinned in environment.yml:28-31, "2.3.1+" per README.md:73), which is absent
from /root/reference and not installable here.  The oracle is therefore pinned
against outputs of the reference's OWN importable modules run in the build
container: oracle/make_golden.py imports /root/reference/models/{loss,model_utils,gnn}.py
verbatim (with oracle/pyg.py's GINEConv injected as `torch_geometric.nn.GINEConv`)
and extracts `build_edge_index_and_attr` from utils/data.py, and writes the
fixtures under tests/golden/.  tests/test_oracle_pinned.py checks every oracle
function against those fixtures.  The one piece with no reference-side pin is
PyG's own GINEConv/Batch arithmetic, restated from its published algorithm:
for that piece parity is "unpinned" (stated again in DESIGN.md).
"""
