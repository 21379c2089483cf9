"""Minimal stand-in for the `pdbx` package (mmcif_pdbx), DEV CONTAINER ONLY: just enough of `PdbxReader` and its data
containers for the UNMODIFIED fr3d.cif.reader (`from pdbx import PdbxReader as Reader`, reader.py:17) to read an mmCIF
file, so that tools/make_golden.py can record what the reference's own reader makes of a CIF text ("add
a ~150-line pdbx shim and use the reference reader unchanged").  Nothing in the product or the tests imports this.

Covers: data_ blocks, key-value categories, loop_ tables, single / double quoted values, ;-delimited text fields,
# comments.  Interface used by the reference: container.name, container.get_object(category) ->
obj.item_name_list ('_category.item'), obj.row_count, obj.row_list."""


class DataCategory(object):
    def __init__(self, name):
        self.name = name
        self.attributes = []
        self.row_list = []

    @property
    def item_name_list(self):
        return ["_%s.%s" % (self.name, a) for a in self.attributes]

    @property
    def row_count(self):
        return len(self.row_list)


class DataContainer(object):
    def __init__(self, name):
        self.name = name
        self._objects = {}

    def get_object(self, name):
        return self._objects.get(name)

    def append(self, obj):
        self._objects[obj.name] = obj


def _tokens(handle):
    """mmCIF tokens of a text: (kind, value), kind in {'data', 'loop', 'name', 'value'}."""
    lines = handle.read().split("\n") if hasattr(handle, "read") else list(handle)
    k = 0
    while k < len(lines):
        line = lines[k]
        if line.startswith(";"):                       # text field up to a line that starts with ';'
            parts = [line[1:]]
            k += 1
            while k < len(lines) and not lines[k].startswith(";"):
                parts.append(lines[k])
                k += 1
            yield "value", "\n".join(parts).strip()
            k += 1
            continue
        i, n = 0, len(line)
        while i < n:
            c = line[i]
            if c in " \t\r":
                i += 1
            elif c == "#":
                break
            elif c in "'\"":                           # quoted: ends at the quote followed by whitespace / end of line
                j = i + 1
                while j < n and not (line[j] == c and (j + 1 == n or line[j + 1] in " \t\r")):
                    j += 1
                yield "value", line[i + 1:j]
                i = j + 1
            else:
                j = i
                while j < n and line[j] not in " \t\r":
                    j += 1
                tok = line[i:j]
                low = tok.lower()
                if low.startswith("data_"):
                    yield "data", tok[5:]
                elif low == "loop_":
                    yield "loop", None
                elif tok.startswith("_"):
                    yield "name", tok
                else:
                    yield "value", tok
                i = j
        k += 1


class PdbxReader(object):
    def __init__(self, handle):
        self.handle = handle

    def read(self, container_list):
        container = None
        toks = list(_tokens(self.handle))
        k = 0
        while k < len(toks):
            kind, val = toks[k]
            if kind == "data":
                container = DataContainer(val)
                container_list.append(container)
                k += 1
            elif kind == "loop":
                k += 1
                names = []
                while k < len(toks) and toks[k][0] == "name":
                    names.append(toks[k][1])
                    k += 1
                cat = names[0][1:].split(".")[0]
                obj = DataCategory(cat)
                obj.attributes = [nm.split(".", 1)[1] for nm in names]
                values = []
                while k < len(toks) and toks[k][0] == "value":
                    values.append(toks[k][1])
                    k += 1
                w = len(names)
                obj.row_list = [values[r:r + w] for r in range(0, len(values) - w + 1, w)]
                container.append(obj)
            elif kind == "name":
                cat, attr = val[1:].split(".", 1)
                obj = container.get_object(cat)
                if obj is None:
                    obj = DataCategory(cat)
                    obj.row_list = [[]]
                    container.append(obj)
                obj.attributes.append(attr)
                obj.row_list[0].append(toks[k + 1][1])
                k += 2
            else:
                k += 1
