"""Static routing tables of Res-ViT (host-side integers): for every position j inside a block of
`block_size` layers, which packed router indices (MSB-first keep bits, res-vit/model.py:169-173)
  - take the low-rank approximator at this layer,
  - run the full transformer layer,
  - are straight-through ("ste") entries.
Values are those produced by the reference's get_indices_from_LRA_mask (res-vit/model_utils.py:69-107) from
its hand-written mapping tables, quirks included; tests compare them with the live
reference function wherever /root/reference exists.
"""

_TABLES = {
    1: [([0], [1], [])],
    2: [([1], [2, 3], [0]), ([0, 2], [1, 3], [])],
    4: [
        ([4, 5, 6, 7], [2, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15], [0, 1, 2, 3]),
        ([2, 3, 10, 11], [2, 4, 5, 6, 7, 10, 12, 13, 14, 15], [0, 1, 8, 9]),
        ([1, 5, 9, 13], [2, 3, 4, 5, 6, 7, 10, 11, 14, 15], [0, 4, 8, 12]),
        ([0, 2, 4, 6, 8, 10, 12, 14], [1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 13, 15], []),
    ],
}


def get_indices_from_LRA_mask(block_size, mapping_table=None):
    if mapping_table is not None:
        raise NotImplementedError("custom mapping tables are not supported; block_size 1, 2 and 4 are built in")
    if block_size not in _TABLES:
        raise ValueError("unsupported block_size %r (the reference supports 1, 2 and 4)" % (block_size,))
    return [(list(a), list(t), list(s)) for a, t, s in _TABLES[block_size]]


def repeat_kv(x, n_rep):
    """GQA head expansion (res-vit/model_utils.py:3-12): identity when n_rep == 1."""
    if n_rep == 1:
        return x
    bs, slen, n_kv, hd = x.shape
    return x[:, :, :, None, :].expand(bs, slen, n_kv, n_rep, hd).reshape(bs, slen, n_kv * n_rep, hd)
