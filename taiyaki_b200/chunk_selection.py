"""Chunk sampling and filtering for training batches -- the batching surface of
taiyaki/chunk_selection.py (same names and semantics: FILTER_PARAMETERS :9-26,
sample_chunks :29-95, sample_filter_parameters :98-123)."""
from collections import defaultdict, namedtuple

import numpy as np

from .maths import med_mad


class FILTER_PARAMETERS(namedtuple('FILTER_PARAMETERS', (
        'filter_mean_dwell', 'filter_max_dwell', 'filter_min_pass_fraction',
        'median_meandwell', 'mad_meandwell', 'model_stride', 'path_buffer'))):
    """Parameters to filter signal chunk selections (chunk_selection.py:9-26)."""


def sample_chunks(read_data, number_to_sample, chunk_len, filter_params,
                  chunk_len_means_sequence_len=False, standardize=True,
                  select_strands_randomly=True, first_strand_index=0):
    """Sample chunks from reads until enough pass the filters; returns
    (chunks, rejection_reason_counts)."""
    nreads = len(read_data)
    number_to_sample_used = nreads if not number_to_sample else number_to_sample
    maximum_attempts_allowed = int(
        number_to_sample_used / filter_params.filter_min_pass_fraction)
    chunks = []
    rejection_reasons = defaultdict(lambda: 0)
    attempts = 0
    while len(chunks) < number_to_sample_used and attempts < maximum_attempts_allowed:
        read_number = (np.random.randint(nreads) if select_strands_randomly else
                       (first_strand_index + attempts) % nreads)
        attempts += 1
        read = read_data[read_number]
        if chunk_len_means_sequence_len:
            raise NotImplementedError('sequence-length chunks are not on the training path')
        chunk = read.get_chunk_with_sample_length(chunk_len, standardize=standardize)
        chunk.apply_filters(filter_params)
        rejection_reasons[chunk.reject_reason] += 1
        if chunk.accepted:
            chunks.append(chunk)
    return chunks, rejection_reasons


def sample_filter_parameters(read_data, number_to_sample, chunk_len, filter_mean_dwell,
                             filter_max_dwell, filter_min_pass_fraction, model_stride,
                             path_buffer, chunk_len_means_sequence_len=False):
    """Median / MAD of the mean dwell over a sample of chunks."""
    no_filter_params = FILTER_PARAMETERS(
        filter_mean_dwell=filter_mean_dwell, filter_max_dwell=filter_max_dwell,
        filter_min_pass_fraction=filter_min_pass_fraction, median_meandwell=None,
        mad_meandwell=None, model_stride=None, path_buffer=None)
    chunks, _ = sample_chunks(read_data, number_to_sample, chunk_len, no_filter_params,
                              chunk_len_means_sequence_len=chunk_len_means_sequence_len)
    meandwells = [chunk.mean_dwell for chunk in chunks]
    median_meandwell, mad_meandwell = med_mad(meandwells)
    return FILTER_PARAMETERS(
        filter_mean_dwell=filter_mean_dwell, filter_max_dwell=filter_max_dwell,
        filter_min_pass_fraction=filter_min_pass_fraction,
        median_meandwell=median_meandwell, mad_meandwell=mad_meandwell,
        model_stride=model_stride, path_buffer=path_buffer)
