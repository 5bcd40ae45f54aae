"""The reference's on-disk data formats, read into flat arrays the GPU path can be fed from.

The reference opens the session CSV and its index files per SAMPLE (one ``open`` + ``seek`` + ``csv.reader`` per
``__getitem__``: data/base/reader.py:60-74, data/datasets/index.py:52-58) and tokenises strings on every access.  A B200
consumes ~50 k tokens per millisecond-scale step, so here every file is read ONCE:

* ``*.session.idx``   pairs of 8-byte native-endian unsigned (start, end) byte offsets into the CSV, count in the last 8
  bytes (data/base/reader.py:18-74, written by data/base/csv_index_builder.py:27-56)        -> (n, 2) int64
* ``*.{loo,nextitem,...}.idx``   pairs (session index, target position), count in the trailer
  (data/datasets/index.py:26-58, index_builder.py:24-49)                                     -> (n, 2) int64
* vocabulary ``token<TAB>id`` (core/tokenization/vocabulary.py:72-91), popularity (one float per id per line,
  init/factories/metrics/sampled_metrics.py:15-17), item-id lists (core/utils/ioutils.py:142-150)
* the CSV itself (tab separated, header row) -> :class:`TokenisedSessions`: all sessions of one column as ONE int32 token
  array plus (n+1) offsets, storable as two ``.npy`` files and memory-mapped afterwards.

``TokenisedSessions.batch`` then produces the tensors the reference's ``SequencePositionDataset`` (truncate to
``[:pos + 1]``, datasets/sequence_position.py:62-70) + ``TargetExtractorProcessor`` (last element becomes the target,
processors/target_extractor.py:52-77) + ``padded_session_collate`` (keep the LAST max_seq_length items, pad on the right,
optional dynamic padding: data/collate.py:49-108) build sample by sample -- as vectorised numpy index arithmetic.
"""
import csv
import io
import os
import sys
from typing import Dict, Iterable, List, Optional, Sequence, Union

import numpy as np
import torch

INT_BYTE_SIZE = 8          # asme/data/datasets/__init__.py
_U64 = np.dtype("=u8")     # native byte order, as int.to_bytes(..., byteorder=sys.byteorder) writes it

ITEM_SEQ_ENTRY_NAME = "item"
TARGET_ENTRY_NAME = "item.target"
SAMPLE_IDS = "sample_ids"


def _read_pairs(path: Union[str, os.PathLike], what: str) -> np.ndarray:
    raw = np.fromfile(path, dtype=_U64)
    if raw.size == 0:
        raise ValueError(f"{what} {path}: empty file (the trailer with the entry count is missing)")
    n = int(raw[-1])
    if raw.size != 2 * n + 1:
        raise ValueError(f"{what} {path}: trailer says {n} entries, file holds {(raw.size - 1) / 2:g}")
    return raw[:-1].reshape(n, 2).astype(np.int64)


def _write_pairs(path: Union[str, os.PathLike], pairs: np.ndarray) -> None:
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    if (pairs < 0).any():
        raise ValueError("index entries are unsigned")
    out = np.empty(pairs.size + 1, dtype=_U64)
    out[:-1] = pairs.reshape(-1)
    out[-1] = pairs.shape[0]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    out.tofile(path)


def read_session_index(path) -> np.ndarray:
    """(n, 2) int64 byte offsets [start, end) of every session in the CSV (CsvDatasetIndex, data/base/reader.py:18-74)"""
    return _read_pairs(path, "session index")


def read_position_index(path) -> np.ndarray:
    """(n, 2) int64 (session index, target position) (SequencePositionIndex, data/datasets/index.py:26-58)"""
    return _read_pairs(path, "position index")


def write_position_index(path, pairs) -> None:
    """the file SequencePositionIndexBuilder.build writes (index_builder.py:24-49)"""
    _write_pairs(path, pairs)


def build_session_index(csv_path, index_path, session_keys: Sequence[str], delimiter: str = "\t") -> np.ndarray:
    """What CsvSessionIndexer.create (data/base/csv_index_builder.py:27-56) produces: consecutive lines with equal values in
    the ``session_keys`` columns form one session; boundaries are byte offsets of whole lines.  One pass over the file
    instead of a peek + re-parse per line.  Returns the (n, 2) boundaries it wrote."""
    with open(csv_path, "rb") as f:
        header_line = f.readline()
        headers = [h.strip() for h in next(csv.reader(io.StringIO(header_line.decode("utf-8")), delimiter=delimiter))]
        missing = [k for k in session_keys if k not in headers]
        if missing:
            raise Exception(f"not all session keys [{list(session_keys)}] could be found in headers [{headers}]")
        cols = [headers.index(k) for k in session_keys]
        bounds: List[List[int]] = []
        current = None
        pos = f.tell()
        for line in f:
            row = next(csv.reader(io.StringIO(line.decode("utf-8")), delimiter=delimiter))
            key = tuple(row[c] for c in cols)
            if key != current:
                bounds.append([pos, pos])
                current = key
            pos += len(line)
            bounds[-1][1] = pos
    pairs = np.asarray(bounds, dtype=np.int64).reshape(-1, 2)
    _write_pairs(index_path, pairs)
    return pairs


def read_vocabulary(path, delimiter: str = "\t") -> Dict[str, int]:
    """token -> id in file order (CSVVocabularyReaderWriter.read, core/tokenization/vocabulary.py:84-91)"""
    with open(path, newline="") as f:
        return {token: int(i) for token, i in csv.reader(f, delimiter=delimiter)}


def read_popularity(path) -> np.ndarray:
    """one float per vocabulary id per line (init/factories/metrics/sampled_metrics.py:15-17)"""
    with open(path) as f:
        return np.asarray([float(line) for line in f if line.strip()], dtype=np.float64)


def read_item_ids(path) -> List[int]:
    """one id per line (load_file_with_item_ids, core/utils/ioutils.py:142-150)"""
    with open(path) as f:
        return [int(line) for line in f if line.strip()]


class TokenisedSessions:
    """All sessions of one CSV column as token ids: ``tokens`` int32 (total), ``offsets`` int64 (n + 1)."""

    def __init__(self, tokens: np.ndarray, offsets: np.ndarray):
        self.tokens, self.offsets = tokens, offsets
        if offsets.ndim != 1 or offsets.size < 1 or int(offsets[-1]) != tokens.shape[0]:
            raise ValueError("offsets do not describe the token array")

    def __len__(self) -> int:
        return self.offsets.size - 1

    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets)

    def sequence(self, i: int) -> np.ndarray:
        return np.asarray(self.tokens[self.offsets[i]:self.offsets[i + 1]])

    @classmethod
    def from_csv(cls, csv_path, session_index: np.ndarray, vocabulary: Dict[str, int], column: str = "item_id",
                 delimiter: str = "\t", unk_token: str = "<UNK>") -> "TokenisedSessions":
        """reads [start, end) of every session (CsvDatasetReader.get_sequence, data/base/reader.py:101-123), takes ``column`` of
        every line (ItemSessionParser.parse, datasets/sequence.py:104-126) and maps tokens to ids, unknown tokens to the id of
        ``unk_token`` (Tokenizer._convert_item_to_id)"""
        with open(csv_path, "rb") as f:
            data = f.read()
        header = [h.strip() for h in next(csv.reader(io.StringIO(data[:data.index(b"\n")].decode("utf-8")), delimiter=delimiter))]
        if column not in header:
            raise KeyError(f"column {column!r} not in the header {header}")
        col = header.index(column)
        unk = vocabulary.get(unk_token)
        tokens: List[int] = []
        offsets = [0]
        for start, end in np.asarray(session_index, dtype=np.int64):
            for row in csv.reader(io.StringIO(data[start:end].decode("utf-8")), delimiter=delimiter):
                tid = vocabulary.get(row[col], unk)
                if tid is None:
                    raise KeyError(f"token {row[col]!r} is not in the vocabulary and there is no {unk_token!r} entry")
                tokens.append(tid)
            offsets.append(len(tokens))
        return cls(np.asarray(tokens, dtype=np.int32), np.asarray(offsets, dtype=np.int64))

    def save(self, prefix) -> None:
        np.save(f"{prefix}.tokens.npy", self.tokens)
        np.save(f"{prefix}.offsets.npy", self.offsets)

    @classmethod
    def load(cls, prefix, mmap: bool = True) -> "TokenisedSessions":
        mode = "r" if mmap else None
        return cls(np.load(f"{prefix}.tokens.npy", mmap_mode=mode), np.load(f"{prefix}.offsets.npy", mmap_mode=mode))

    def batch(self, session_ids: Iterable[int], positions: Optional[Iterable[int]] = None, max_seq_length: int = 200,
              pad_id: int = 0, extract_target: bool = True, dynamic_padding: bool = False,
              pin: bool = False, entry_name: str = ITEM_SEQ_ENTRY_NAME) -> Dict[str, torch.Tensor]:
        """One collated batch: ``item`` (N, S) int64 right-padded, ``item.target`` (N) (with ``extract_target``), ``length``,
        ``sample_ids`` and, when ``positions`` are given (a position-index slice), ``pos``.  Per sample the reference computes
        ``seq = session[:pos + 1]``; ``target = seq[-1]``, ``item = seq[:-1]``; ``item = item[-S:]`` + right padding;
        ``length`` is the length BEFORE the cut to S (collate.py:82), as the reference reports it.  ``entry_name``: key of the
        sequence (a store of another CSV column, e.g. an item attribute, collates under its own name and ``<name>.target``)."""
        sid = np.asarray(list(session_ids) if not isinstance(session_ids, np.ndarray) else session_ids, dtype=np.int64)
        if sid.size and (sid.min() < 0 or sid.max() >= len(self)):
            raise Exception(f"{int(sid.max() if sid.max() >= len(self) else sid.min())} is not a valid index in [0, {len(self)}]")
        start = np.asarray(self.offsets)[sid]
        full = np.asarray(self.offsets)[sid + 1] - start
        if positions is not None:
            pos = np.asarray(list(positions) if not isinstance(positions, np.ndarray) else positions, dtype=np.int64)
            n_seq = np.minimum(pos + 1, full)                      # python slicing semantics of [:pos + 1]
        else:
            pos, n_seq = None, full
        n_in = n_seq - 1 if extract_target else n_seq
        if extract_target and (n_seq < 1).any():
            raise IndexError("target extraction needs at least one item per sample")
        S = int(min(max_seq_length, max(int(n_in.max()) if n_in.size else 0, 0))) if dynamic_padding else int(max_seq_length)
        kept = np.minimum(n_in, S)
        first = start + (n_in - kept)                              # keep the LAST S items
        cols = np.arange(S, dtype=np.int64)[None, :]
        valid = cols < kept[:, None]
        src = np.where(valid, first[:, None] + cols, 0)
        item = np.where(valid, np.asarray(self.tokens)[src.reshape(-1)].reshape(src.shape), pad_id).astype(np.int64)
        out = {entry_name: torch.from_numpy(item), "length": torch.from_numpy(n_in.astype(np.int64)),
               SAMPLE_IDS: torch.from_numpy(sid)}
        if extract_target:
            out[entry_name + ".target"] = torch.from_numpy(np.asarray(self.tokens)[start + n_seq - 1].astype(np.int64))
        if pos is not None:
            out["pos"] = torch.from_numpy(pos)
        if pin and torch.cuda.is_available():
            out = {k: v.pin_memory() for k, v in out.items()}
        return out


def epoch_batches(n_entries: int, batch_size: int, shuffle: bool = True, seed: int = 0, epoch: int = 0, world: int = 1, rank: int = 0,
                  drop_last: bool = False) -> List[np.ndarray]:
    """Entry indices of one epoch for this rank, cut into batches -- the order ``torch.utils.data.DistributedSampler`` (what
    Lightning's ``ddp`` puts in front of the reference's datasets) + a ``DataLoader(batch_size=...)`` produce: a permutation
    seeded with ``seed + epoch``, padded to a multiple of ``world`` by wrapping around (or truncated with ``drop_last``), rank r
    takes entries r, r + world, ...  Feed the batches to ``TokenisedSessions.batch(positions[idx, 0], positions[idx, 1], ...)``."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside [0, {world})")
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        order = torch.randperm(n_entries, generator=g).numpy()
    else:
        order = np.arange(n_entries)
    if drop_last and n_entries % world != 0:
        total = (n_entries // world) * world
        order = order[:total]
    else:
        total = -(-n_entries // world) * world
        pad = total - order.size
        if pad > 0:
            reps = -(-pad // max(order.size, 1))
            order = np.concatenate([order, np.tile(order, reps)[:pad]])
    mine = order[rank:total:world]
    return [mine[i:i + batch_size] for i in range(0, mine.size, batch_size)]
