"""Drop-in entry points of fr3d/classifiers/NA_pairwise_interactions.py, computed on a B200.

Same names, argument meaning and return types as the reference:

  annotate_nt_nt_in_structure(structure, categories, focused_basepair_cutoffs={},
                              ideal_hydrogen_bonds={}, chains=[], timerData=None,
                              get_datapoint=False)                         NA:1213-1243
  generatePairwiseAnnotation(entry_id, chain_id, inputPath, output, category, format)   NA:3150-3291
  focus_basepair_cutoffs, reverse_edges, get_parent, myTimer,
  write_txt_output_file, write_ebi_json_output_file, write_unit_data_file

plus the batched form the GPU wants:

  annotate_batch(batch_or_specs, categories, ...) -> list of per-structure result tuples

The per-pair rules run in CUDA (csrc/classify*.cuh) and so does the order-dependent tail
(same-edge cleanup, emission, crossing numbers, covalent links: csrc/tail.cuh); Python only
turns the final integer rows into unit-id tuples, lazily.  postprocess.py is the host form of
the tail (tail="host").  `get_datapoint=True` (rule-development diagnostics incl. hydrogen-bond
checks, NA:2404-2536) is outside the accelerated path and raises.
"""

import argparse
import json
import os
import pickle
from collections import defaultdict
from time import time

import numpy as np

from .. import _lib
from .. import definitions as defs
from .. import packing
from .. import postprocess
from .. import tail as tailmod
from ..engine import BoxTable, Engine, reverse_edges  # noqa: F401  (reverse_edges re-exported)

nt_nt_screen_distance = 12        # NA:69
near_discrepancy_cutoff = 1.0     # NA:71
nt_nt_cutoffs = defs.nt_nt_cutoffs()
load_ideal_basepair_hydrogen_bonds = defs.load_ideal_basepair_hydrogen_bonds
get_parent = defs.get_parent

Leontis_Westhof_basepairs = ['cWW', 'cSS', 'cHH', 'cHS', 'cHW', 'cSH', 'cSW', 'cWH', 'cWS', 'tSS', 'tHH', 'tHS',
                             'tHW', 'tSH', 'tSW', 'tWH', 'tWS', 'tWW']     # NA:3160

_CATEGORY_BITS = {'sO': _lib.CAT_SO, 'stacking': _lib.CAT_STACKING, 'sugar_ribose': _lib.CAT_SUGAR_RIBOSE,
                  'backbone': _lib.CAT_BACKBONE, 'coplanar': _lib.CAT_COPLANAR}


def focus_basepair_cutoffs(basepair_cutoffs, interactions):
    """NA:88-136: keep the families requested, split by the sign of the normal."""
    focused = {}
    lower = set()
    if interactions:
        for interaction in interactions:
            lower.add(interaction.lower())
    else:
        for combination in basepair_cutoffs.keys():
            for interaction in basepair_cutoffs[combination]:
                lower.add(interaction.lower())
    for combination in basepair_cutoffs.keys():
        focused[combination] = {1: {}, -1: {}}
        for interaction in basepair_cutoffs[combination].keys():
            family = interaction.lower()[0:3]
            if family in lower:
                subcat = list(basepair_cutoffs[combination][interaction].keys())[0]
                sign = 1 if basepair_cutoffs[combination][interaction][subcat]["normalmin"] > 0 else -1
                focused[combination][sign][interaction] = dict(basepair_cutoffs[combination][interaction])
    return focused


def myTimer(state, data={}):
    """NA:138-171 (same shared-default-dict behaviour)."""
    if "currentState" in data:
        data[data["currentState"]] += time() - data["lastTime"]
    if state == "summary":
        total = 0.000000000001
        for st in data["allStates"]:
            if not st == "lastTime" and not st == "currentState":
                total += data[st]
        print("Summary of time taken:")
        for st in data["allStates"]:
            if not st == "lastTime" and not st == "currentState":
                print("%-31s: %10.3f seconds %10.3f minutes %10.3f%% of total" % (st, data[st], data[st] / 60,
                                                                              100 * data[st] / total))
        print("%-31s: %10.3f seconds %10.3f minutes %10.3f%% of total" % ("Total", total, total / 60, 100))
    elif state not in data:
        data[state] = 0
        if "allStates" in data:
            data["allStates"].append(state)
        else:
            data["allStates"] = [state]
    data["currentState"] = state
    data["lastTime"] = time()
    return data


# ------------------------------------------------------------------ tables cache --
_box_cache = {}


def _tables(focused, hbonds):
    """(BoxTable, per-box hydrogen-bond atom sets, tail.LabelTable) of a focused cutoff table, cached."""
    key = (id(focused), id(hbonds))
    ent = _box_cache.get(key)
    if ent is None or ent[0] is not focused or ent[1] is not hbonds:
        bt = BoxTable(focused)
        atoms = postprocess.box_atom_sets(bt, bt.box_combo, hbonds)
        ent = (focused, hbonds, bt, atoms, tailmod.LabelTable(bt, atoms))
        if len(_box_cache) > 16:
            _box_cache.clear()
        _box_cache[key] = ent
    return ent[2], ent[3], ent[4]


def _box_table(focused, hbonds):
    return _tables(focused, hbonds)[:2]


_default_focus = {}


def _default_focused(interactions):
    key = tuple(interactions) if interactions else ()
    if key not in _default_focus:
        _default_focus[key] = focus_basepair_cutoffs(nt_nt_cutoffs, list(key))
    return _default_focus[key]


def category_mask(categories):
    m = 0
    for k, bit in _CATEGORY_BITS.items():
        if k in categories:
            m |= bit
    return m


# ------------------------------------------------------------------ batched API --
_TAIL_JOB = None        # (batch, hits, cuts, boxtable, box_atoms, categories) of the call in progress, read by forked workers


def _tail_range(bounds):
    batch, hits, cuts, bt, box_atoms, categories = _TAIL_JOB
    return [postprocess.build_results(batch, s, hits[cuts[s]:cuts[s + 1]], bt, box_atoms, categories)
            for s in range(bounds[0], bounds[1])]


def _host_tail(batch, hits, cuts, bt, box_atoms, categories, workers):
    """postprocess.build_results for every structure.  The tail is pure Python / numpy and independent per
    structure (SURVEY §8 A8 keeps it on the host), so with workers > 1 it runs in forked worker processes: they
    inherit the batch and the hit records copy-on-write and send back the result dicts.  The children never touch
    CUDA."""
    global _TAIL_JOB
    S = batch.n_structures
    if workers is None:
        workers = int(os.environ.get("FR3D_TAIL_WORKERS", "1"))
    if workers == 0:
        workers = os.cpu_count() or 1
    workers = max(1, min(workers, S // 8))
    _TAIL_JOB = (batch, hits, cuts, bt, box_atoms, categories)
    try:
        if workers == 1:
            return _tail_range((0, S))
        import multiprocessing
        n_jobs = min(S, workers * 4)                 # a few jobs per worker: structures differ in size
        edges = [int(round(k * S / n_jobs)) for k in range(n_jobs + 1)]
        with multiprocessing.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_tail_range, list(zip(edges[:-1], edges[1:])), chunksize=1)
        return [r for part in parts for r in part]
    finally:
        _TAIL_JOB = None


def annotate_batch(batch, categories, focused_basepair_cutoffs=None, ideal_hydrogen_bonds=None, device=None,
                   pinned=None, return_raw=False, tail_workers=None, tail="device", as_arrays=False):
    """Annotate every structure of an NtBatch (or of a list of coordinate specs) in one pass:
    one H2D copy, K1 (if frames are not given), K2/K3, K4, the order-dependent tail, one D2H copy.

    tail="device" (default): same-edge cleanup, emission, crossing numbers and covalent links run on the GPU
    (fr3d_annotation_tail); the result is a tail.BatchAnnotations - a sequence whose item s is the reference's
    (interaction_to_list_of_tuples, category_to_interactions) of structure s, built from the row arrays on first
    access.  as_arrays=True returns the same object; it says the caller will read .rows_of(s) / .rows (label id,
    nt1, nt2, crossing; .labels.name(id)) and never needs the tuples.
    tail="host": the round-1 path - hit records back, postprocess.build_results per structure in Python
    (tail_workers: processes for it; None = $FR3D_TAIL_WORKERS or 1, 0 = one per host core); returns a list.
    With return_raw also the extras dict (sorted hit array, candidate pairs, frames)."""
    if not isinstance(batch, packing.NtBatch):
        batch = packing.pack_specs(batch)
    if not focused_basepair_cutoffs:
        focused_basepair_cutoffs = _default_focused(categories.get('basepair', []))
    if not ideal_hydrogen_bonds:
        ideal_hydrogen_bonds = load_ideal_basepair_hydrogen_bonds()
    bt, box_atoms, labels = _tables(focused_basepair_cutoffs, ideal_hydrogen_bonds)
    eng = Engine.get(device)
    if tail == "device":
        hits, n_pairs, extras = eng.annotate_batch(batch, bt, category_mask(categories), pinned=pinned,
                                                   return_frames=return_raw, labels=labels, want_hits=return_raw)
        results = tailmod.BatchAnnotations(batch, extras["rows"], extras["row_start"], extras["row_count"], labels,
                                           n_pairs=n_pairs)
        if return_raw:
            extras["n_pairs"] = n_pairs
            extras["hits"] = postprocess.sort_hits(hits)
            return results, extras
        return results
    if tail != "host":
        raise ValueError("tail must be 'device' or 'host'")
    if as_arrays:
        raise ValueError("as_arrays needs tail='device'")
    hits, n_pairs, extras = eng.annotate_batch(batch, bt, category_mask(categories), pinned=pinned,
                                               return_frames=return_raw)
    hits = postprocess.sort_hits(hits)
    # hits are sorted by rank = first_seen * 16 + stencil and first_seen is a global nt index of the
    # same structure as nt1, so the sorted array is already grouped by structure, in order
    sidx = np.searchsorted(batch.struct_off, hits['i'], side='right') - 1
    cuts = np.searchsorted(sidx, np.arange(batch.n_structures + 1))
    results = _host_tail(batch, hits, cuts, bt, box_atoms, categories, tail_workers)
    if return_raw:
        extras["n_pairs"] = n_pairs
        extras["hits"] = hits
        return results, extras
    return results


def annotate_nt_nt_in_structure(structure, categories, focused_basepair_cutoffs={}, ideal_hydrogen_bonds={},
                                chains=[], timerData=None, get_datapoint=False):
    """NA:1213-1243.  `structure` is an fr3d Structure (the reference's own class or
    fr3d_python_b200.data.Structure).  Returns (interaction_to_list_of_tuples,
    category_to_interactions, timerData, pair_to_data)."""
    if get_datapoint:
        raise NotImplementedError("get_datapoint=True (diagnostic mode with hydrogen-bond checks, NA:2404-2536) "
                                  "is outside the GPU path; use the reference for rule development")
    if not focused_basepair_cutoffs:
        focused_basepair_cutoffs = _default_focused(categories['basepair'])
    if not ideal_hydrogen_bonds:
        ideal_hydrogen_bonds = load_ideal_basepair_hydrogen_bonds()
    if not timerData:
        timerData = myTimer("start")
    timerData = myTimer("Packing nucleotides", timerData)
    soa = getattr(structure, "_soa_batch", None)
    if soa is not None and not chains:
        batch = soa
    else:
        batch = packing.pack_structures(structure, chains)
    timerData = myTimer("Annotating interactions", timerData)
    results = annotate_batch(batch, categories, focused_basepair_cutoffs, ideal_hydrogen_bonds)
    interaction_to_list_of_tuples, category_to_interactions = results[0]
    timerData = myTimer("Calculate crossing", timerData)
    pair_to_data = defaultdict(dict)
    return interaction_to_list_of_tuples, category_to_interactions, timerData, pair_to_data


# ------------------------------------------------------------------ writers --
def write_txt_output_file(outputNAPairwiseInteractions, pdbid, interaction_to_list_of_tuples, categories,
                          category_to_interactions):
    """NA:3057-3093: one file per category, rows 'u1 \\t label \\t u2 \\t crossing' sorted by
    (model, chain, int(number), label)."""
    for category in categories:
        if category == "near":
            continue
        filename = os.path.join(outputNAPairwiseInteractions, pdbid + "_" + category + ".txt")
        quads = []
        for interaction in sorted(category_to_interactions[category]):
            if category == 'basepair':
                inter = interaction.replace("w", "W").replace("s", "S").replace("h", "H")
            else:
                inter = interaction
            if len(categories[category]) == 0 or inter in categories[category] or \
                    ("near" in categories and "n" in interaction and inter.replace('n', '') in categories[category]):
                for a, b, c in interaction_to_list_of_tuples[interaction]:
                    quads.append((a, inter, b, c))
        ordered = sorted(quads, key=lambda x: (x[0].split("|")[1], x[0].split("|")[2], int(x[0].split("|")[4]), x[1]))
        with open(filename, 'w') as f:
            for a, b, c, d in ordered:
                f.write("%s\t%s\t%s\t%s\n" % (a, b, c, d))


def write_ebi_json_output_file(outputNAPairwiseInteractions, pdbid, interaction_to_list_of_tuples, categories,
                               category_to_interactions, chain, unit_id_to_sequence_position, modified):
    """NA:3095-3146."""
    for category in categories.keys():
        filename = os.path.join(outputNAPairwiseInteractions, pdbid + "_" + chain + "_" + category + ".json")
        output = {"pdb_id": pdbid, "chain_id": chain, "modified": modified}
        annotations = []
        for interaction in category_to_interactions[category]:
            inter = interaction
            if category == 'basepair':
                inter = interaction.replace("w", "W").replace("s", "S").replace("h", "H")
            if len(categories[category]) == 0 or inter in categories[category]:
                for a, b, c in interaction_to_list_of_tuples[interaction]:
                    f1 = a.split("|")
                    f2 = b.split("|")
                    if f1[2] == chain and f2[2] == chain:
                        if unit_id_to_sequence_position[a] < unit_id_to_sequence_position[b]:
                            annotations.append({
                                "seq_id1": str(unit_id_to_sequence_position[a]), "3d_id1": f1[4], "nt1": f1[3],
                                "unit1": f1[3], "bp": inter, "seq_id2": str(unit_id_to_sequence_position[b]),
                                "nt2": f2[3], "unit2": f2[3], "3d_id2": f2[4], "crossing": str(c)})
        output["annotations"] = annotations
        with open(filename, 'w') as f:
            f.write(json.dumps(output))


def write_unit_data_file(PDB, unit_data_path, structure):
    """NA:3000-3054: per model/chain pickle [units, order, glycosidic centres, rotations], protocol 2."""
    if len(unit_data_path) > 0:
        nucleotides = structure.residues(type=["RNA linking", "DNA linking"])
        all_nts = {}
        for nt in nucleotides:
            fields = nt.unit_id().split("|")
            ident = "_".join(fields[0:3])
            symmetry = fields[8] if len(fields) == 9 else ""
            all_nts.setdefault(ident, []).append((symmetry, nt.index, nt))
        for ident in all_nts.keys():
            filename = os.path.join(unit_data_path, "units", ident + "_NA.pickle")
            units, order, cntrs, rttns = [], [], [], []
            for symmetry, index, nt in sorted(all_nts[ident], key=lambda p: (p[0], p[1])):
                units.append(nt.unit_id())
                order.append(nt.index)
                cntrs.append(nt.centers["glycosidic"])
                rttns.append(nt.rotation_matrix)
            with open(filename, 'wb') as fh:
                pickle.dump([units, order, cntrs, rttns], fh, 2)
            print("  Wrote unit data file %s" % filename)


# ------------------------------------------------------------------ driver --
def load_structure(filename, pdbid=""):
    """NA:174-294 without the download step: CIF parsing stays on the host and belongs to the
    reference's fr3d.cif.reader (out of scope here).  Accepts the JSON
    coordinate specs this package uses (.json / .json.gz), or defers to fr3d.cif.reader if it
    is importable."""
    from ..data import Structure
    message = []
    for cand in (filename, filename + ".json.gz", filename + ".json", filename + ".cif", filename + ".cif.gz"):
        if os.path.exists(cand):
            filename = cand
            break
    else:
        return None, ["Code is not clever enough to find %s" % filename]
    try:
        if filename.endswith(".json") or filename.endswith(".json.gz"):
            import gzip
            opener = gzip.open if filename.endswith(".gz") else open
            with opener(filename, "rt") as f:
                spec = json.load(f)
            structure = Structure.from_spec(spec)
        else:
            import gzip
            from fr3d.cif.reader import Cif      # reference host code, if installed
            opener = gzip.open if filename.endswith(".gz") else open
            with opener(filename, "rt") as raw:
                structure = Cif(raw).structure()
        message.append("Loaded " + filename)
        return structure, message
    except Exception as ex:      # same contract as the reference: (None, messages)
        message.append("Could not load %s due to exception %s: %s" % (filename, type(ex).__name__, ex))
        return None, message


def generatePairwiseAnnotation(entry_id, chain_id, inputPath, outputNAPairwiseInteractions, category, output_format):
    """NA:3150-3291."""
    if isinstance(entry_id, str):
        entry_id = [entry_id]
    categories = {}
    if category:
        for cat in category.split(","):
            categories[cat] = []
    else:
        categories['basepair'] = Leontis_Westhof_basepairs
    if 'basepair_detail' in categories:
        categories['basepair'] = Leontis_Westhof_basepairs + ['cWB', 'cBW']
    elif 'basepair' in categories:
        categories['basepair'] = Leontis_Westhof_basepairs
    else:
        categories['basepair'] = ['cWW']
    if len(inputPath) > 0 and not os.path.exists(inputPath):
        print("Attempting to create input path %s" % inputPath)
        os.mkdir(inputPath)
    if len(outputNAPairwiseInteractions) > 0 and not os.path.exists(outputNAPairwiseInteractions):
        print("Attempting to create output path %s" % outputNAPairwiseInteractions)
        os.mkdir(outputNAPairwiseInteractions)
    PDBs = []
    for entry in entry_id:
        path_split = os.path.split(entry)
        if len(path_split[0]) > 0:
            PDBs.append(path_split)
        else:
            PDBs.append((inputPath, entry))
    timerData = myTimer("start")
    failed_structures = []
    counter = 0
    if chain_id:
        if len(entry_id) > 1:
            print("chain argument can only be used with a single PDB file")
            PDBs = []
        else:
            chains = chain_id.split(",")
    else:
        chains = []
    focused = focus_basepair_cutoffs(nt_nt_cutoffs, categories['basepair'])
    hbonds = load_ideal_basepair_hydrogen_bonds()
    for path, PDB in PDBs:
        counter += 1
        pdbid = PDB.replace(".cif", "").replace(".pdb", "").replace(".json", "").replace(".gz", "")
        filename = os.path.join(path, PDB)
        print("Reading file %s, which is number %d out of %d" % (filename, counter, len(PDBs)))
        timerData = myTimer("Reading CIF files", timerData)
        structure, messages = load_structure(filename, pdbid)
        if not structure:
            for message in messages:
                failed_structures.append((pdbid, message))
            continue
        i2t, c2i, timerData, pair_to_data = annotate_nt_nt_in_structure(structure, categories, focused, hbonds, chains,
                                                                       timerData)
        timerData = myTimer("Recording interactions", timerData)
        print("  Recording interactions in %s" % outputNAPairwiseInteractions)
        if output_format == 'txt':
            write_txt_output_file(outputNAPairwiseInteractions, pdbid, i2t, categories, c2i)
        elif output_format == 'ebi_json':
            if chains:
                bases = structure.residues(chain=chains, type=["RNA linking", "DNA linking"])
            else:
                bases = structure.residues(type=["RNA linking", "DNA linking"])
            pos = {}
            modified = {}
            for base in bases:
                chain = base.chain
                if chain not in pos:
                    pos[chain] = {}
                    modified[chain] = []
                pos[chain][base.unit_id()] = base.index
                fields = base.unit_id().split('|')
                if fields[3] not in ['A', 'C', 'G', 'U', 'DA', 'DC', 'DG', 'DT']:
                    modified[chain].append({'seq_id': str(base.index), 'nt1': fields[3], 'unit1': fields[3],
                                            '3d_id': fields[4]})
            for chain in list(pos.keys()):
                write_ebi_json_output_file(outputNAPairwiseInteractions, pdbid, i2t, categories, c2i, chain,
                                           pos[chain], modified[chain])
        else:
            print('Output format %s not recognized' % output_format)
    myTimer("summary", timerData)
    if len(failed_structures) > 0:
        print("Error messages:")
        for message in failed_structures:
            print("%s %s" % message)
    else:
        print("All files read successfully")


def main(argv=None):
    """CLI of NA:3293-3337."""
    parser = argparse.ArgumentParser()
    parser.add_argument('PDBfiles', type=str, nargs='+', help='.cif / .json filename(s)')
    parser.add_argument('-o', "--output", help="Output Location of Pairwise Interactions")
    parser.add_argument('-i', "--input", help='Input Path')
    parser.add_argument('-c', "--category",
                        help='Interaction category or categories (basepair,stacking,sO,backbone,coplanar,'
                             'basepair_detail,covalent,sugar_ribose,near)')
    parser.add_argument('-f', "--format", help='Output format (txt,ebi_json)')
    parser.add_argument("--chain", help='Chain or chains separated by commas, no spaces; only for one PDB file')
    args = parser.parse_args(argv)
    generatePairwiseAnnotation(args.PDBfiles, args.chain or None, args.input or "", args.output or "",
                               args.category or 'basepair', args.format or 'txt')


if __name__ == "__main__":
    main()
