"""cProfile of the host tail of annotate_batch (hit records -> reference-format tuples) on the GPU box."""
import cProfile
import pstats
import sys

sys.path.insert(0, ".")
from fr3d_python_b200 import synthetic
from fr3d_python_b200.classifiers import NA_pairwise_interactions as NA

sys.path.insert(0, "tests")
from conftest import ALL9

base = synthetic.pack_base()
batch = synthetic.make_structure_batch(base, 100, seed0=4000)
NA.annotate_batch(batch, ALL9)
pr = cProfile.Profile()
pr.enable()
NA.annotate_batch(batch, ALL9)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
