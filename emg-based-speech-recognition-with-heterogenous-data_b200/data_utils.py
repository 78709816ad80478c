"""Host-side boundary helpers, same names and semantics as the reference's data_utils.py:165-185."""
import torch

PAD = 42


def combine_fixed_length(tensor_list, length):
    """data_utils.py:165-174: concatenate utterances along time, pad the tail with the VALUE 42 to a multiple of `length`,
    view as (n, length, channels)."""
    total_length = sum(t.size(0) for t in tensor_list)
    if total_length % length != 0:
        pad_length = length - (total_length % length)
        tensor_list = list(tensor_list)
        tensor_list.append(torch.full((pad_length, *tensor_list[0].size()[1:]), PAD, dtype=tensor_list[0].dtype,
                                      device=tensor_list[0].device))
        total_length += pad_length
    tensor = torch.cat(tensor_list, 0)
    n = total_length // length
    return tensor.view(n, length, *tensor.size()[1:])


def decollate_tensor(tensor, lengths):
    """data_utils.py:176-185 (host mirror; on the device the same gather is sst_gather_rows_pad)."""
    b, s, d = tensor.size()
    tensor = tensor.view(b * s, d)
    results = []
    idx = 0
    for length in lengths:
        assert idx + length <= b * s
        results.append(tensor[idx:idx + length])
        idx += length
    return results
