"""FASTA/FASTQ input and bit-packing of aligned sequences.

`fasta2dic` mirrors apples/fasta2dic.py:42-72 (same alphabet normalisation): names are the header up to the
first blank; sequences are upper-cased (or lower-case masked to '-' with mask_flag); letters outside the
alphabet become '-' (nucleotide: everything but A,C,G,T; protein: B,J,O,U,X,Z).

The packers turn the resulting byte rows into the device layout (DESIGN.md "data layout"):
  nucleotide: three bit-planes per sequence (lo, hi, valid), 32 sites per uint32 word, site s -> bit s%32 of word s//32
              code A=0 C=1 G=2 T=3, lo = code&1, hi = code>>1, valid = (byte != '-'); gap sites have lo=hi=0
  protein:    one uint8 code per site: 0..19 in the reference's a2i order (distance.py:418-678, A R N D C Q E G H
              I L K M F P S T W Y V), 20 for '-'; any other byte maps to 0 ('A') exactly as a2i's NA=0 does
"""
import numpy as np

NUC = 0
AA = 1

_AA_ORDER = 'ARNDCQEGHILKMFPSTWYV'
AA_GAP = 20


def _records(fp):
    """FASTA/FASTQ reader with the record semantics of fasta2dic.py:4-39 (header char '>' or '@', name up to the
    first blank, multi-line sequences, optional '+' quality block)."""
    name = None
    seqs = []
    it = iter(fp)
    for line in it:
        if not line:
            continue
        c = line[0]
        if c in '>@':
            if name is not None:
                yield name, ''.join(seqs)
            name = line[1:].rstrip('\r\n').partition(' ')[0]
            seqs = []
        elif c == '+' and name is not None:
            # fastq quality block: skip as many characters as the sequence has
            seq = ''.join(seqs)
            need = len(seq)
            got = 0
            while got < need:
                try:
                    q = next(it)
                except StopIteration:
                    break
                got += len(q.rstrip('\r\n'))
            yield name, seq
            name = None
            seqs = []
        elif name is not None:
            seqs.append(line.rstrip('\r\n'))
    if name is not None:
        yield name, ''.join(seqs)


def fasta2dic(ref_fp, prot_flag, mask_flag):
    """{name: np.ndarray(dtype='S1')}, insertion-ordered, same normalisation as fasta2dic.py:42-72."""
    refs = {}
    mask_translation = str.maketrans('abcdefghijklmnopqrstuvwxyz', '-' * 26)
    if prot_flag:
        invalid_translation = str.maketrans('BJOUXZ', '-' * 6)
    else:
        invalid_translation = str.maketrans('BDEFHIJKLMNOPQRSUVWXYZ', '-' * 22)
    with open(ref_fp) as f:
        for name, seq in _records(f):
            s = seq.translate(mask_translation) if mask_flag else seq.upper()
            refs[name] = np.frombuffer(s.translate(invalid_translation).encode(), dtype='S1')
    return refs


class FastaMatrix:
    """An alignment file read by the native reader (libapples_b200: apples_fasta_open, the C++ twin of fasta2dic above):
    `names` (list of str, file order) and `matrix` (uint8 [n, L] view of the reader's buffer -- pinned host memory when a
    CUDA device is usable, so the placement call DMAs straight from it).  Keep the object alive while `matrix` is used."""

    def __init__(self, path, prot_flag=False, mask_flag=False, threads=0, pinned=True):
        import ctypes as C
        from . import _lib
        self._lib = _lib.load()
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self._lib.apples_fasta_open(str(path).encode(), 1 if prot_flag else 0, 1 if mask_flag else 0, int(threads),
                                         1 if pinned else 0, C.byref(h), err, 512)
        if rc != 0:
            raise OSError('reading %s failed: %s' % (path, err.value.decode(errors='replace')))
        self._h = h
        lib = self._lib
        self.n = int(lib.apples_fasta_count(h))
        self.L = int(lib.apples_fasta_max_len(h))
        self.stride = int(lib.apples_fasta_stride(h))
        self.uniform = bool(lib.apples_fasta_uniform(h))
        self.pinned = bool(lib.apples_fasta_pinned(h))
        n = self.n
        self.lengths = np.ctypeslib.as_array(C.cast(lib.apples_fasta_lengths(h), C.POINTER(C.c_int64)), shape=(n,)) if n else np.zeros(0, np.int64)
        self._name_off = np.ctypeslib.as_array(C.cast(lib.apples_fasta_name_offsets(h), C.POINTER(C.c_int64)), shape=(n + 1,))
        self._name_bytes = C.string_at(lib.apples_fasta_names(h), int(self._name_off[n])) if n else b''
        full = np.ctypeslib.as_array(C.cast(lib.apples_fasta_matrix(h), C.POINTER(C.c_uint8)), shape=(max(n, 1), max(self.stride, 1)))
        self.full = full[:n]                 # [n, stride], rows padded with '-'
        self.matrix = self.full[:, :self.L]  # [n, L] (row stride = self.stride)
        self._names = None

    @property
    def names(self):
        if self._names is None:
            # text files are read as UTF-8 by the Python twin (open() default); invalid bytes raise like there
            self._names = self._name_bytes.decode('utf-8').split('\0')[:-1] if self.n else []
        return self._names

    def close(self):
        if getattr(self, '_h', None):
            self.matrix = self.full = self.lengths = self._name_off = None
            self._lib.apples_fasta_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def as_dict(self, copy=True):
        """{name: 'S1' row} like fasta2dic (later duplicates overwrite earlier ones in place, like a dict).  copy=False
        returns views into the reader's buffer: keep this object alive as long as the dict is used."""
        out = {}
        lens = self.lengths.tolist()
        for i, nm in enumerate(self.names):
            row = self.matrix[i, :lens[i]]
            out[nm] = (row.copy() if copy else row).view('S1')
        return out


def read_alignment(path, prot_flag=False, mask_flag=False, threads=0, pinned=True):
    """Product-path reader: the native C++ twin of fasta2dic (apples_fasta_open).  Fails loudly without the library."""
    return FastaMatrix(path, prot_flag, mask_flag, threads, pinned)


def words_per_row(L):
    """uint32 words per bit-plane row, padded to a multiple of 4 words (16 bytes) so rows can be moved with
    16-byte bulk copies."""
    return ((int(L) + 31) // 32 + 3) // 4 * 4


def as_byte_matrix(seqs, L=None):
    """list/iterable of 'S1' arrays (or a 2-D uint8/S1 array) -> contiguous uint8 [n, L]."""
    if isinstance(seqs, np.ndarray) and seqs.ndim == 2:
        return np.ascontiguousarray(seqs.view(np.uint8))
    seqs = list(seqs)
    if L is None:
        L = len(seqs[0]) if seqs else 0
    out = np.empty((len(seqs), L), dtype=np.uint8)
    for i, s in enumerate(seqs):
        if len(s) != L:
            raise ValueError('sequence %d has length %d, alignment has %d columns' % (i, len(s), L))
        out[i] = np.frombuffer(s, dtype=np.uint8) if not isinstance(s, np.ndarray) else s.view(np.uint8)
    return out


_NUC_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b'ACGT'):
    _NUC_CODE[_c] = _i
_NUC_CODE[ord('-')] = 4


def pack_nucleotide(mat):
    """uint8 [n, L] of bytes in {A,C,G,T,-} -> uint32 [n, 3, W] (planes lo, hi, valid; W = words_per_row(L)).

    The reference compares raw bytes (distance.py:733-737), so a byte that survives fasta2dic but is not one of
    A,C,G,T,- (a non-letter such as '.', '*', '?', a digit) would act as a fifth symbol there.  The 2-bit code
    cannot express that, so such input is rejected loudly instead of silently changing results.
    """
    mat = np.ascontiguousarray(mat, dtype=np.uint8)
    n, L = mat.shape
    W = words_per_row(L)
    code = _NUC_CODE[mat]
    if code.size and code.max() == 255:
        bad = np.unique(mat[code == 255])
        raise ValueError('nucleotide alignment contains bytes the 2-bit packing cannot express: %r'
                         % [chr(b) for b in bad.tolist()])
    planes = np.zeros((n, 3, W * 4), dtype=np.uint8)
    valid = code < 4
    for k, bits in enumerate(((code & 1).astype(bool) & valid, ((code >> 1) & 1).astype(bool) & valid, valid)):
        pb = np.packbits(bits, axis=1, bitorder='little')
        planes[:, k, :pb.shape[1]] = pb
    return planes.view('<u4').reshape(n, 3, W)


_AA_CODE = np.zeros(256, dtype=np.uint8)  # NA = 0: unknown bytes count as 'A' (distance.py:418)
for _i, _c in enumerate(_AA_ORDER):
    _AA_CODE[ord(_c)] = _i
    _AA_CODE[ord(_c.lower())] = _i
_AA_CODE[ord('-')] = AA_GAP


def aa_row_bytes(L):
    """bytes per protein code row, padded to a multiple of 16."""
    return (int(L) + 15) // 16 * 16


def pack_protein(mat):
    """uint8 [n, L] of bytes -> uint8 [n, aa_row_bytes(L)] of codes 0..19, gap=20 (padding columns are gaps)."""
    mat = np.ascontiguousarray(mat, dtype=np.uint8)
    n, L = mat.shape
    out = np.full((n, aa_row_bytes(L)), AA_GAP, dtype=np.uint8)
    out[:, :L] = _AA_CODE[mat]
    return out


def pack(mat, kind):
    return pack_protein(mat) if kind == AA else pack_nucleotide(mat)
