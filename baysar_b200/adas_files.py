"""ADAS ADF11 / ADF15 text files -> the ``(block, ne, Te)`` tables the posterior's constructor consumes.

This is the data format on the input side of the posterior path: the reference reads
hydrogen photon-emissivity coefficients and ionisation / recombination / radiated-power rate coefficients from OpenADAS
text files (`/root/reference/OpenADAS/read_adf15.py:26-110`, `/root/reference/OpenADAS/read_adf11.py:20-69`).  The
parsers below restate those two readers (same tolerance to the files' layout, same array orientation, same ``10**`` on
ADF11), :func:`format_adf11` / :func:`format_adf15` write tables back in the same layout (used to produce the synthetic
fixtures, since no ADAS file may be redistributed or downloaded here), and :class:`FileADAS` is an
:class:`~baysar_b200.atomic_data.AtomicDataProvider` over files: the hydrogen tables as the reference loads them, and the
impurity primitives the reference takes from the proprietary ADAS library (``run_adas406`` / old-signature readers,
`plasmas.py:983-1022,1215-1216`) rebuilt from the element's ADF11 / ADF15 files.

Parity: `tests/test_adas_files.py` compares the parsers with the reference's own functions on committed synthetic
files (`oracle/make_adf_golden.py`, arrays in `tests/golden/adf_parsed.npz`).
"""
import math
import pathlib

import numpy as np

from .atomic_data import AtomicDataProvider, GridTable

ADF15_SEPARATOR = "C-----------------------------------------------------------------------"
_ADF11_BLOCK_START = "---------------------/ "
_ADF15_BLOCK_WORDS = ("EXCIT", "RECOM", "CHEXC", "TYPE", "type")


def _text(source):
    """File path or the text itself (the reference's ``passed=True``)."""
    if isinstance(source, pathlib.Path) or (isinstance(source, str) and "\n" not in source):
        with open(source, "r") as f:
            return f.read()
    return source


def parse_adf11(source):
    """ADF11 text -> GridTable(block 1..n, ne [cm-3], Te [eV], rates [cm3/s or W cm3]).

    Follows `OpenADAS/read_adf11.py:20-69`: the first three integers of line 1 are (blocks, n_ne, n_Te); the log10
    grids are every number between the last ``---`` before the first block line and that line (the remainder of the
    dashed line is the token the reference drops); each block is ``ceil(n_ne n_Te / 8)`` lines of log10 rates stored
    Te-major, returned as ``[ne, Te]``; grids and rates go through ``10**``.
    """
    lines = _text(source).split("\n")
    starts = [i for i, line in enumerate(lines) if line.startswith(_ADF11_BLOCK_START)]
    if not starts:
        raise ValueError("Could not find the first block of rates in the adf11 file!")
    start = starts[0]
    n_blocks, n_ne, n_te = (int(tok) for tok in lines[0].split()[:3])
    flat = " ".join(lines[:start]).split("---")[-1].split()[1:]
    flat = np.array([float(tok) for tok in flat])
    ne, te = flat[:n_ne], flat[n_ne:]
    if len(te) != n_te:
        raise ValueError(f"adf11 grid has {len(flat)} values, header says {n_ne} + {n_te}")
    len_block = int(math.ceil(n_ne * n_te / 8))
    rates = []
    for b in range(n_blocks):
        lhs = 1 + start + b * (1 + len_block)
        vals = np.array([float(tok) for tok in " ".join(lines[lhs:lhs + len_block]).split()])
        rates.append(vals.reshape((n_te, n_ne)).T)
    return GridTable(1 + np.arange(n_blocks), np.power(10, ne), np.power(10, te), np.power(10, np.array(rates)))


def parse_adf15(source):
    """ADF15 text -> GridTable(block 1..n, ne [cm-3], Te [eV], PEC [cm3/s]).

    Follows `OpenADAS/read_adf15.py:20-110`: line 1 starts with the number of blocks; the data section runs from the
    first line naming a transition type (exactly two of EXCIT / RECOM / CHEXC / TYPE / type) to the first comment
    separator and splits into equally long blocks; a block header carries (n_ne, n_Te) after the wavelength (one token
    further right when the wavelength unit ``A`` is a separate token), followed by the ne grid, the Te grid and the
    coefficients stored ne-major.  The axes of the LAST block label the table (as in the reference).
    """
    raw = _text(source)
    all_lines = raw.split("\n")
    n_blocks = int(all_lines[0].split()[0])
    first = None
    for i, line in enumerate(all_lines):
        if sum(word in line for word in _ADF15_BLOCK_WORDS) == 2:
            first = i
            break
    if first is None:
        raise ValueError("No line found!")
    data_lines = raw.split(ADF15_SEPARATOR)[0].split("\n")[first:-1]
    if len(data_lines) % n_blocks:
        raise ValueError("adf15 data section does not split into equal blocks")
    per = len(data_lines) // n_blocks
    tables, ne, te = [], None, None
    for b in range(n_blocks):
        block = data_lines[b * per:(b + 1) * per]
        header = block[0].split()
        shift = 1 if header[1] == "A" else 0
        n_ne, n_te = int(header[1 + shift]), int(header[2 + shift])
        vals = np.array([float(tok) for tok in " ".join(block[1:]).split()])
        ne, te = vals[:n_ne], vals[n_ne:n_ne + n_te]
        tables.append(vals[n_ne + n_te:].reshape((n_ne, n_te)))
    return GridTable(1 + np.arange(n_blocks), ne, te, np.array(tables))


def _rows_of_8(values, fmt):
    values = list(values)
    return ["".join(fmt % v for v in values[i:i + 8]) for i in range(0, len(values), 8)]


def format_adf11(table, element="HYDROGEN", adf11type="scd", comment="synthetic table (baysar_b200), not ADAS data"):
    """Write a GridTable in the ADF11 layout :func:`parse_adf11` (and the reference reader) understands: log10 grids and
    rates with five decimals, eight numbers per line."""
    n_b, n_ne, n_te = table.data.shape
    out = [f"{n_b:5d}{n_ne:5d}{n_te:5d}{1:5d}{n_b:5d}     /{element:<19s}/{adf11type.upper()} SYNTHETIC"]
    out.append("-" * 80)
    out += _rows_of_8(np.log10(table.ne), "%10.5f")
    out += _rows_of_8(np.log10(table.te), "%10.5f")
    for b in range(n_b):
        out.append(f"---------------------/ IPRT= 1  / IGRD= 1  /--------/ Z1={b + 1:2d}   / DATE= 20/10/26")
        out += _rows_of_8(np.log10(table.data[b]).T.ravel(), "%10.5f")
    out.append("C" + "-" * 79)
    out.append("C  " + comment)
    out.append("C" + "-" * 79)
    return "\n".join(out) + "\n"


def format_adf15(table, wavelengths=None, types=None, label="H 0", comment="synthetic table (baysar_b200), not ADAS data"):
    """Write a GridTable in the ADF15 layout :func:`parse_adf15` (and the reference reader) understands: one block per
    transition with its own copy of the grids, coefficients in ``%9.2E`` (the precision of real ADF15 files), eight
    numbers per line."""
    n_b, n_ne, n_te = table.data.shape
    out = [f"{n_b:5d}    /{label} SYNTHETIC PHOTON EMISSIVITY COEFFICIENTS/"]
    for b in range(n_b):
        wl = wavelengths[b] if wavelengths is not None else 4000.0 + 10.0 * b
        typ = types[b] if types is not None else ("EXCIT" if b < n_b * 6 // 10 else "RECOM")
        out.append(f"{wl:9.1f} A{n_ne:5d}{n_te:5d} /FILMEM = synth   /TYPE = {typ} /INDM = T/ISEL ={b + 1:4d}")
        out += _rows_of_8(table.ne, "%9.2E")
        out += _rows_of_8(table.te, "%9.2E")
        for i in range(n_ne):
            out += _rows_of_8(table.data[b, i], "%9.2E")
    out.append(ADF15_SEPARATOR)
    out.append("C  " + comment)
    out.append(ADF15_SEPARATOR)
    return "\n".join(out) + "\n"


def _loglog_interp(table, ne_grid, te_grid, dens, te):
    """``table[ne, Te]`` (positive) at ``dens x te``: bilinear in ``(log10 ne, log10 Te)`` of ``log10(table)``, linear
    extrapolation from the edge cells outside the grid (the rule `scipy.interpolate.interpn(..., fill_value=None)` and the
    reference's ``DataArray.interp`` use for the hydrogen tables, `plasmas.py:723-733`) -> ``[len(dens), len(te)]``."""
    from scipy.interpolate import interpn

    pts = np.stack(np.meshgrid(np.log10(np.asarray(dens, dtype=float)), np.log10(np.asarray(te, dtype=float)),
                               indexing="ij"), axis=-1)
    out = interpn((np.log10(ne_grid), np.log10(te_grid)), np.log10(table), pts, method="linear", bounds_error=False,
                  fill_value=None)
    return np.power(10.0, out)


class FileADAS(AtomicDataProvider):
    """Atomic data from ADAS text files.

    ``adf15``: path or text of the hydrogen PEC file (`plasmas.py:1290-1291` loads ``get_adf15("h", 0, year=12)``);
    ``adf11``: mapping ``{"scd": path, "acd": path, "plt": path, "prb": path}`` for hydrogen (`plasmas.py:706-721`).

    Impurities (`plasmas.py:939-1081, 1207-1243`): ``impurity_adf11 = {"N": {"scd": ..., "acd": ..., "plt": ..., "prb":
    ...}, ...}`` and ``impurity_adf15 = {pec key: path or text}``, the key being the ``pec`` entry of the line catalogue
    (the full string or its basename).  With them this provider supplies the three primitives the reference takes from
    the proprietary ADAS Python library -- ``run_adas406`` (time-dependent ionisation balance: the matrix exponential
    of the SCD / ACD rate matrix, `baysar_b200.ionisation_balance.time_dependent_abundance`), ``read_adf11`` and
    ``read_adf15`` (tables interpolated onto the requested ``(Te, ne)`` grid in log-log space) -- so a posterior with
    impurity lines is built from files alone.  ``impurities`` (another provider) is only consulted for what no file
    was given for.
    """

    def __init__(self, adf15, adf11, impurities=None, impurity_adf11=None, impurity_adf15=None):
        self._adf15 = parse_adf15(adf15)
        self._adf11 = {key: parse_adf11(src) for key, src in adf11.items()}
        self._impurities = impurities
        self._imp11 = {elem: {k: parse_adf11(src) for k, src in files.items()} for elem, files in (impurity_adf11 or {}).items()}
        self._imp15_src = dict(impurity_adf15 or {})
        self._imp15 = {}
        self._balance_cache = {}

    def hydrogen_adf15(self, element="h", charge=0):
        return self._adf15

    def hydrogen_adf11(self, adf11type):
        if adf11type not in self._adf11:
            raise KeyError(f"no adf11 file given for {adf11type!r}")
        t = self._adf11[adf11type]
        return GridTable(t.block[:1], t.ne, t.te, t.data[:1])  # the reference selects block 1 (`plasmas.py:723`)

    def _delegate(self, what):
        if self._impurities is None:
            raise NotImplementedError(f"no file given for {what} and no fallback provider (impurities=...)")
        return self._impurities

    def _on_grid(self, elem, kind, te, dens):
        """All blocks of an element's adf11 table interpolated to ``[block, len(dens), len(te)]``."""
        t = self._imp11[elem][kind]
        data = np.array([_loglog_interp(t.data[b], t.ne, t.te, dens, te) for b in range(len(t.block))])
        return GridTable(t.block, np.asarray(dens, dtype=float), np.asarray(te, dtype=float), data)

    def ionisation_balance(self, elem, te, dens, tint, meta):
        if elem not in self._imp11 or not {"scd", "acd"} <= set(self._imp11[elem]):
            return self._delegate(f"the scd/acd tables of {elem}").ionisation_balance(elem, te, dens, tint, meta)
        from .ionisation_balance import time_dependent_abundance

        key = (elem, np.asarray(te, dtype=float).tobytes(), np.asarray(dens, dtype=float).tobytes())
        if key not in self._balance_cache:  # the constructor asks for 30 times on one grid
            self._balance_cache = {key: (self._on_grid(elem, "scd", te, dens), self._on_grid(elem, "acd", te, dens))}
        scd, acd = self._balance_cache[key]
        return time_dependent_abundance(scd, acd, tint, meta)

    def adf11_rates(self, elem, adf11type, is1, te, dens):
        if elem not in self._imp11 or adf11type not in self._imp11[elem]:
            return self._delegate(f"the {adf11type} table of {elem}").adf11_rates(elem, adf11type, is1, te, dens)
        t = self._imp11[elem][adf11type]
        return _loglog_interp(t.select(is1), t.ne, t.te, dens, te).T  # [len(te), len(dens)] like the old read_adf11

    def _pec_table(self, file):
        if file not in self._imp15:
            src = self._imp15_src.get(file, self._imp15_src.get(pathlib.PurePosixPath(str(file)).name))
            self._imp15[file] = parse_adf15(src) if src is not None else None
        return self._imp15[file]

    def adf15_pecs(self, file, block, te, dens):
        t = self._pec_table(file)
        if t is None:
            return self._delegate(f"the adf15 file {file!r}").adf15_pecs(file, block, te, dens)
        return _loglog_interp(t.select(block), t.ne, t.te, dens, te).T  # [len(te), len(dens)]
