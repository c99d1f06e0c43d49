"""Shared test helpers: image readers, random corpora, result comparison (src/lib.rs:46-66 rule)."""
from __future__ import annotations

import random
from typing import List, Sequence, Tuple

import numpy as np


def image_arrays(im) -> dict:
    """numpy views of a pb_index_image (host pointers owned by the builder)."""
    as_arr = np.ctypeslib.as_array
    F = im.num_fields
    nn, ne, nt, nrp, nd = int(im.n_nodes), int(im.n_edges), int(im.n_terms), int(im.n_rows_padded), int(im.n_docs)
    out = {
        "node_edge_begin": as_arr(im.node_edge_begin, shape=(nn + 1,)),
        "node_term_lo": as_arr(im.node_term_lo, shape=(nn,)),
        "node_term_hi": as_arr(im.node_term_hi, shape=(nn,)),
        "node_parent": as_arr(im.node_parent, shape=(nn,)),
        "node_char": as_arr(im.node_char, shape=(nn,)),
        "edge_char": as_arr(im.edge_char, shape=(ne,)) if ne else np.zeros(0, np.uint32),
        "edge_child": as_arr(im.edge_child, shape=(ne,)) if ne else np.zeros(0, np.uint32),
        "term_row_begin": as_arr(im.term_row_begin, shape=(nt + 1,)),
        "term_byte_len": as_arr(im.term_byte_len, shape=(nt,)) if nt else np.zeros(0, np.uint32),
        "term_node": as_arr(im.term_node, shape=(nt,)) if nt else np.zeros(0, np.uint32),
    }
    # tile-blocked columns [tile][1 + 2F][128] -> plain per-row columns
    blocks = as_arr(im.post_blocks, shape=(nrp // 128, 1 + 2 * F, 128))
    out.update({
        "post_doc": blocks[:, 0, :].reshape(-1),
        "post_tf": [blocks[:, 1 + f, :].reshape(-1) for f in range(F)],
        "post_fl": [blocks[:, 1 + F + f, :].reshape(-1) for f in range(F)],
        "doc_key": as_arr(im.doc_key, shape=(nd,)) if nd else np.zeros(0, np.uint64),
        "removed": as_arr(im.removed_bitmap, shape=((nd + 31) // 32 + 1,)),
    })
    return out


def image_term_string(a: dict, t: int) -> str:
    cps = []
    n = int(a["term_node"][t])
    while n != 0:
        cps.append(chr(int(a["node_char"][n])))
        n = int(a["node_parent"][n])
    return "".join(reversed(cps))


def image_find_node(a: dict, term: str):
    node = 0
    for ch in term:
        lo, hi = int(a["node_edge_begin"][node]), int(a["node_edge_begin"][node + 1])
        chars = a["edge_char"][lo:hi]
        i = int(np.searchsorted(chars, ord(ch)))
        if i >= len(chars) or int(chars[i]) != ord(ch):
            return None
        node = int(a["edge_child"][lo + i])
    return node


def image_expand(a: dict, term: str) -> List[str]:
    node = image_find_node(a, term)
    if node is None:
        return []
    return [image_term_string(a, t) for t in range(int(a["node_term_lo"][node]), int(a["node_term_hi"][node]))]


WORDS = ["a", "ab", "abc", "abd", "abcd", "abcde", "b", "ba", "bab", "c", "ca", "xyz", "xy", "x", "q", "zz",
         "zzz", "héllo", "hé", "日本", "日本語", "oy", "oysters", "the", "the,", "then"]


def random_corpus(rng: random.Random, n_docs: int, n_fields: int, max_len: int = 6, multi_value: bool = False):
    """docs: [(key, [[values of field 0], [values of field 1], ...])] with ' '-separated tokens;
    sometimes double spaces (empty tokens) and empty fields."""
    docs = []
    for key in range(n_docs):
        fields = []
        for _ in range(n_fields):
            nvals = rng.choice([1, 1, 1, 2, 0]) if multi_value else 1
            vals = []
            for _ in range(nvals):
                n = rng.randint(0, max_len)
                toks = [rng.choice(WORDS) for _ in range(n)]
                sep = "  " if rng.random() < 0.1 else " "
                vals.append(sep.join(toks))
            fields.append(vals)
        docs.append((key, fields))
    return docs


def random_query(rng: random.Random) -> str:
    n = rng.randint(1, 4)
    toks = []
    for _ in range(n):
        w = rng.choice(WORDS)
        if rng.random() < 0.5:
            w = w[: rng.randint(1, len(w))]
        if rng.random() < 0.1:
            w = "nomatch"
        toks.append(w)
    sep = "  " if rng.random() < 0.15 else " "
    return sep.join(toks)


def assert_same_results(got: Sequence[Tuple[int, float]], exp: Sequence[Tuple[int, float]], ctx="", tol=1e-9,
                        exact=True):
    """Comparison rule of src/lib.rs:46-66: both sides sorted by (score desc, key asc); key sets
    bit-exact; scores within `tol` (and bit-exact when `exact`)."""
    g = sorted(got, key=lambda r: (-r[1], r[0]))
    e = sorted(exp, key=lambda r: (-r[1], r[0]))
    assert sorted(k for k, _ in g) == sorted(k for k, _ in e), f"{ctx}: doc-id sets differ\n got={g}\n exp={e}"
    gd, ed = dict(g), dict(e)
    for k in gd:
        assert abs(gd[k] - ed[k]) <= tol, f"{ctx}: key {k}: {gd[k]!r} vs {ed[k]!r}"
        if exact:
            assert gd[k] == ed[k], f"{ctx}: key {k}: {gd[k]!r} != {ed[k]!r} (not bit-exact)"
    if exact:
        assert g == e, f"{ctx}: order differs\n got={g}\n exp={e}"
