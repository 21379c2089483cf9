"""Host data model with the interface of fr3d.data (Atom, Component, Structure) that the hot
path touches (fr3d/data/atoms.py:9, components.py:118-202, structures.py:13-70, base.py:80-209).

Only what annotate_nt_nt_in_structure, the writers and the geometry shims read is provided.
Base frames are NOT fitted on the host: Component / Structure call the CUDA frame-fit kernel
(K1) — Structure.calculate_rotation_matrix() fits every residue in one launch.  Hydrogen
placement (infer_NA_hydrogens) stays host arithmetic on top of the fitted frame, as in the
reference (components.py:452-462).
"""

import numpy as np

from .. import definitions as defs
from ..packing import encode_unit_id


class Atom(object):
    """fr3d/data/atoms.py:9-48."""

    def __init__(self, pdb=None, model=None, chain=None, component_id=None, component_number=None,
                 component_index=None, insertion_code=None, alt_id=None, x=None, y=None, z=None, group=None,
                 type=None, name=None, symmetry=None, polymeric=None):
        self.pdb = pdb
        self.model = model
        self.chain = chain
        self.component_id = component_id
        self.component_number = component_number
        self.component_index = component_index
        self.insertion_code = insertion_code
        self.alt_id = alt_id
        self.x = x
        self.y = y
        self.z = z
        self.group = group
        self.type = type
        self.name = name
        self.symmetry = symmetry
        self.polymeric = polymeric

    def coordinates(self):
        return np.array([self.x, self.y, self.z])

    def __repr__(self):
        return '<Atom: %s %s>' % (self.name, self.coordinates())


class Centers(object):
    """AtomProxy (fr3d/data/base.py:80-209): name -> coordinates of the atom(s) with that name
    (mean if several, empty array if none), plus explicitly set / defined centers."""

    def __init__(self, atoms):
        self._atoms = atoms
        self._data = {}
        self._definitions = {}

    def define(self, name, atoms):
        self._definitions[name] = atoms
        self._data[name] = set([atoms]) if isinstance(atoms, str) else set(atoms)

    def setcenter(self, name, vector):
        self._data[name] = vector
        self._definitions[name] = ["Base atoms"]

    def definition(self, name):
        return self._definitions.get(name)

    def _coordinates(self, names):
        coords = [a.coordinates() for a in self._atoms if a.name in names]
        if not coords:
            return np.array(coords)
        if len(coords) == 1:
            return coords[0]
        return np.average(coords, axis=0)

    def __getitem__(self, key):
        if isinstance(key, (list, set, tuple)):
            return self._coordinates(set(key))
        if key not in self._data:
            return self._coordinates(set([key]))
        if isinstance(self._data[key], set):
            self._data[key] = self._coordinates(self._data[key])
        return self._data.get(key, np.array([]))

    def __contains__(self, key):
        return key in self._data or any(a.name == key for a in self._atoms)


def _select(items, kwargs):
    """EntitySelector filtering (fr3d/data/base.py:21-75): attribute == value, attribute in
    collection, or callable(attribute)."""
    out = []
    for obj in items:
        ok = True
        for key, value in kwargs.items():
            attr = getattr(obj, key, None)
            if callable(attr):
                attr = attr()
            if callable(value):
                ok = bool(value(attr))
            elif isinstance(value, (list, set, tuple)):
                ok = attr in set(value)
            else:
                ok = attr == value
            if not ok:
                break
        if ok:
            out.append(obj)
    return out


class Component(object):
    """fr3d/data/components.py:118-202 for nucleotides."""

    def __init__(self, atoms, pdb=None, model=None, type=None, chain=None, symmetry=None, sequence=None,
                 number=None, index=None, insertion_code=None, polymeric=None, alt_id=None, defer_frame=False):
        self._atoms = atoms
        self.pdb = pdb
        self.model = model
        self.type = type
        self.chain = chain
        self.symmetry = symmetry
        self.sequence = sequence
        self.number = number
        self.index = index
        self.insertion_code = insertion_code
        self.polymeric = polymeric
        self.alt_id = alt_id
        self.base_center = None
        self.rotation_matrix = None
        self.centers = Centers(self._atoms)
        if not defer_frame:
            self.calculate_rotation_matrix()
            self.infer_NA_hydrogens()
            self._define_centers()

    # -- frame: one-residue launch of the batched CUDA kernel (components.py:273-349)
    def calculate_rotation_matrix(self):
        from ..geometry.frames import fit_component_frames
        fit_component_frames([self])

    def _set_frame(self, U, center):
        self.rotation_matrix = np.matrix(U)          # numpy.matrix like the reference
        self.base_center = np.array(center)

    def infer_NA_hydrogens(self):
        """components.py:359-462 for standard residues: fix amino-hydrogen labels, then add every
        missing base hydrogen at centre + U . h_std."""
        if self.rotation_matrix is None:
            return
        if self.sequence not in defs.NAbasehydrogens:
            mod = defs.modified_table().get(self.sequence)
            if mod is None or any('H' in a.name for a in self._atoms):
                return
            # components.py:505-513: no hydrogen observed -> an ideal-position copy of EVERY mapped base
            # atom (heavy atoms included, mapping.py:76-79) is appended under the modified residue's names
            std = defs.NAbasecoordinates[mod["parent"]]
            U = np.asarray(self.rotation_matrix)
            for name in mod["base_atoms"]:
                new = self.base_center + U.dot(np.array(std[mod["modified_atom_to_parent"][name]]))
                self._atoms.append(Atom(name=name, x=float(new[0]), y=float(new[1]), z=float(new[2])))
            return
        parent = defs.SEQ_TO_PARENT[self.sequence]
        already = set(a.name for a in self._atoms)
        amino = {'A': ("N7", "H61", "H62"), 'C': ("C5", "H41", "H42"), 'G': ("N1", "H21", "H22")}.get(parent)
        if amino:
            by = {a.name: a for a in self._atoms}
            if all(n in by for n in amino):
                hv, h1, h2 = (by[n] for n in amino)
                d1 = (hv.x - h1.x) ** 2 + (hv.y - h1.y) ** 2 + (hv.z - h1.z) ** 2
                d2 = (hv.x - h2.x) ** 2 + (hv.y - h2.y) ** 2 + (hv.z - h2.z) ** 2
                if d1 < d2:
                    (h1.x, h1.y, h1.z), (h2.x, h2.y, h2.z) = (h2.x, h2.y, h2.z), (h1.x, h1.y, h1.z)
        coords = defs.NAbasecoordinates[self.sequence]
        U = np.asarray(self.rotation_matrix)
        for h in defs.NAbasehydrogens[self.sequence]:
            if h not in already:
                new = self.base_center + U.dot(np.array(coords[h]))
                self._atoms.append(Atom(name=h, x=float(new[0]), y=float(new[1]), z=float(new[2])))

    def _define_centers(self):
        """components.py:162-179."""
        self.centers = Centers(self._atoms)
        if self.base_center is not None:
            self.centers.setcenter('base', self.base_center)
            parent = defs.get_parent(self.sequence)
            if self.sequence in defs.SEQ_TO_PARENT:
                self.centers.define('glycosidic', ['N9'] if parent in ('A', 'G') else ['N1'])
            elif parent is not None:
                p2m = defs.modified_table()[self.sequence]["parent_atom_to_modified"]
                want = 'N9' if parent in ('A', 'G') else 'N1'
                self.centers.define('glycosidic', [p2m.get(want, want)])

    def atoms(self, **kwargs):
        name = kwargs.get('name')
        if isinstance(name, str):
            definition = self.centers.definition(name)
            if definition:
                kwargs['name'] = definition
        return _select(self._atoms, kwargs)

    def coordinates(self, **kwargs):
        return np.array([atom.coordinates() for atom in self.atoms(**kwargs)])

    def unit_id(self):
        return encode_unit_id(self.pdb, self.model, self.chain, self.sequence, self.number, self.alt_id,
                              self.insertion_code, self.symmetry)

    def __repr__(self):
        return '<Component %s>' % self.unit_id()


class Structure(object):
    """fr3d/data/structures.py:13-70."""

    def __init__(self, residues, pdb=None, model=None, chain=None, symmetry=None, breaks=None, operators=None):
        self.pdb = pdb
        self.model = model
        self.chain = chain
        self.symmetry = symmetry
        self._residues = residues
        self._breaks = breaks
        self._operators = operators or {}
        self._soa_batch = None

    def residues(self, **kwargs):
        if 'polymeric' not in kwargs:
            kwargs['polymeric'] = True
        if kwargs.get('polymeric', False) is None:
            kwargs.pop('polymeric')
        return _select(self._residues, kwargs)

    def calculate_rotation_matrix(self):
        """structures.py:58-63, but ONE launch for all residues."""
        from ..geometry.frames import fit_component_frames
        fit_component_frames(self._residues)

    def infer_NA_hydrogens(self):
        for residue in self._residues:
            residue.infer_NA_hydrogens()
            residue._define_centers()

    def __len__(self):
        return len(self._residues)

    @classmethod
    def from_spec(cls, spec):
        """Build a Structure from a plain coordinate spec: components are created with the frame
        fit deferred, all frames are fitted in one K1 launch, hydrogens are then placed on the host
        (SURVEY §8(b) integration note)."""
        comps = []
        for nt in spec["nts"]:
            seq = nt["sequence"]
            atoms = [Atom(pdb=spec["pdb"], model=nt["model"], chain=nt["chain"], component_id=seq,
                          component_number=nt["number"], component_index=nt["index"],
                          insertion_code=nt.get("insertion_code"), alt_id=nt.get("alt_id"), x=a[1], y=a[2], z=a[3],
                          name=a[0], symmetry=nt["symmetry"], polymeric=True, type=a[0][0]) for a in nt["atoms"]]
            ctype = "DNA linking" if seq in ("DA", "DC", "DG", "DT") else "RNA linking"
            comps.append(Component(atoms, pdb=spec["pdb"], model=nt["model"], type=ctype, chain=nt["chain"],
                                   symmetry=nt["symmetry"], sequence=seq, number=nt["number"], index=nt["index"],
                                   insertion_code=nt.get("insertion_code"), alt_id=nt.get("alt_id"),
                                   polymeric=True, defer_frame=True))
        st = cls(comps, pdb=spec["pdb"])
        st.calculate_rotation_matrix()
        st.infer_NA_hydrogens()
        return st
