"""CPU: the numpy restatements of the reference-side ops around the path (oracle/oracle.py):
StringSplit and the three pre-lookup filter ops."""
import numpy as np

from oracle import oracle as O


def test_string_split_restatement_matches_tf_doc_examples():
    """string_split_op.cc semantics on the cases of TF's own docs/tests (string_split_op_test.py):
    'hello world' / 'a b c' on ' ' and skip_empty on '#a' ',b' ..."""
    toks, rs = O.string_split([b"hello world", b"a b c"], b" ")
    assert toks == [b"hello", b"world", b"a", b"b", b"c"] and rs.tolist() == [0, 2, 5]
    toks, rs = O.string_split([b"#a", b"b#", b"#c#"], b"#", skip_empty=False)
    assert toks == [b"", b"a", b"b", b"", b"", b"c", b""] and rs.tolist() == [0, 2, 4, 7]
    toks, rs = O.string_split([b"#a", b"b#", b"#c#"], b"#", skip_empty=True)
    assert toks == [b"a", b"b", b"c"] and rs.tolist() == [0, 1, 2, 3]
    toks, rs = O.string_split([b"hello", b""], b"")
    assert toks == [b"h", b"e", b"l", b"l", b"o"] and rs.tolist() == [0, 5, 5]


def test_filter_ops_restatement():
    """gather_indice_value_ops.cc:33-70, select_value_ops.cc:33-56,
    gather_str_value_gen_indice_ops.cc:31-58 with the interval test l <= x <= r."""
    idx = np.array([[0, 0], [0, 1], [2, 0], [3, 0]], np.int64)
    vals = np.array([-1, 5, 9, 70], np.int64)
    i2, v2 = O.gather_indice_value(idx, vals, [(0, 9)])
    assert v2.tolist() == [5, 9] and i2.tolist() == [[0, 1], [2, 0]]
    assert O.select_value(vals, [(0, 9), (60, 80)], 3).tolist() == [3, 5, 9, 70]
    i3, s3 = O.gather_str_value_gen_indice([b"a", b"", b"bc", b""], b"", dense_rank=2)
    assert s3 == [b"a", b"bc"] and i3.tolist() == [[0, 0], [2, 0]]
