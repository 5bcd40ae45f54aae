"""Golden files for the on-disk formats, produced by the UNMODIFIED reference classes in the build container:

    python tests/golden/make_formats_golden.py

Writes tests/golden/formats/:
  sessions.csv                  synthetic sessions (seeded), tab separated with header, two session-key columns
  sessions.session.idx          written by the reference's CsvSessionIndexer        (data/base/csv_index_builder.py)
  sessions.vocabulary.item_id.txt   written by the reference's CSVVocabularyReaderWriter (core/tokenization/vocabulary.py)
  sessions.nextitem.idx         written by the reference's SequencePositionIndexBuilder (data/datasets/index_builder.py)
  sessions.popularity.item_id.txt   one float per id
  expected.json                 what the reference's readers / dataset / processor / collate return for these files:
                                session boundaries, position pairs, vocabulary, tokenised sessions, and collated batches
                                (SequencePositionDataset + TargetExtractorProcessor + padded_session_collate; plain dataset
                                without target extraction; dynamic padding)
This script is the only code that touches /root/reference; tests read the committed files.
"""
import json
import os
import random
import sys
from pathlib import Path

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shims"))
import ref_shims  # noqa: E402

ref_shims.install()

from asme.core.tokenization.tokenizer import Tokenizer  # noqa: E402
from asme.core.tokenization.vocabulary import CSVVocabularyReaderWriter, VocabularyBuilder  # noqa: E402
from asme.data.base.csv_index_builder import CsvSessionIndexer  # noqa: E402
from asme.data.base.reader import CsvDatasetIndex, CsvDatasetReader  # noqa: E402
from asme.data.collate import PadDirection, PadInformation, padded_session_collate  # noqa: E402
from asme.data.datasets.index import SequencePositionIndex  # noqa: E402
from asme.data.datasets.index_builder import SequencePositionIndexBuilder  # noqa: E402
from asme.data.datasets.processors.target_extractor import TargetExtractorProcessor  # noqa: E402
from asme.data.datasets.processors.tokenizer import TokenizerProcessor  # noqa: E402
from asme.data.datasets.sequence import ItemSequenceDataset, ItemSessionParser, MetaInformation, PlainSequenceDataset  # noqa: E402
from asme.data.datasets.sequence_position import SequencePositionDataset  # noqa: E402
from asme.data.utils.csv import create_indexed_header, read_csv_header  # noqa: E402

OUT = Path(HERE) / "formats"
OUT.mkdir(exist_ok=True)
rng = random.Random(22)

# ---- the CSV: 23 sessions of 1..17 events, item names with spaces / unicode, a token that will be missing from the vocabulary
items = [f"Item {i}" for i in range(1, 41)] + ["Ünïcode item", "item,with,commas"]
rows = ["session_id\tuser_id\titem_id\trating"]
n_events = []
for s in range(23):
    n = rng.choice([1, 2, 3, 5, 8, 13, 17])
    n_events.append(n)
    for _ in range(n):
        rows.append(f"{s // 2}\t{s % 2}\t{rng.choice(items)}\t{rng.randint(1, 5)}")
rows[5] = rows[5].rsplit("\t", 2)[0] + "\tNever seen item\t3"
csv_path = OUT / "sessions.csv"
csv_path.write_text("\n".join(rows) + "\n", encoding="utf-8")

# ---- files written by the reference
index_path = OUT / "sessions.session.idx"
CsvSessionIndexer(delimiter="\t").create(csv_path, index_path, ["session_id", "user_id"])

vb = VocabularyBuilder()
for t in ["<PAD>", "<MASK>", "<UNK>"] + items:
    vb.add_token(t)
vocabulary = vb.build()
vocab_path = OUT / "sessions.vocabulary.item_id.txt"
with vocab_path.open("w", newline="") as f:
    CSVVocabularyReaderWriter().write(vocabulary, f)
tokenizer = Tokenizer(vocabulary, pad_token="<PAD>", mask_token="<MASK>", unk_token="<UNK>")

pop_path = OUT / "sessions.popularity.item_id.txt"
pop = np.random.default_rng(5).random(len(vocabulary))
pop[:3] = 0.0
pop_path.write_text("\n".join(repr(float(x)) for x in pop / pop.sum()) + "\n")

features = [MetaInformation("item", "str", tokenizer=tokenizer, column_name="item_id")]
header = create_indexed_header(read_csv_header(csv_path, "\t"))
reader = CsvDatasetReader(csv_path, CsvDatasetIndex(index_path))
plain = PlainSequenceDataset(reader, ItemSessionParser(header, features, delimiter="\t"))
tok = TokenizerProcessor({"item": tokenizer})
dataset = ItemSequenceDataset(plain, processors=[tok])
pos_path = OUT / "sessions.nextitem.idx"
SequencePositionIndexBuilder().build(dataset, pos_path)        # every position 1 .. len-1 of every session
pos_index = SequencePositionIndex(pos_path)

# ---- what the reference reads back
idx = CsvDatasetIndex(index_path)
expected = {"num_sessions": len(idx), "boundaries": [[idx.get(i).start, idx.get(i).end] for i in range(len(idx))],
            "positions": [list(pos_index[i]) for i in range(len(pos_index))],
            "vocabulary": {t: vocabulary.get_id(t) for t in vocabulary.tokens()},
            "sessions": [dataset[i]["item"] for i in range(len(dataset))]}
with vocab_path.open() as f:
    again = CSVVocabularyReaderWriter().read(f)
assert [again.get_id(t) for t in again.tokens()] == [vocabulary.get_id(t) for t in vocabulary.tokens()]


def tolist(batch):
    return {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in batch.items()}


def collate(samples, max_len, dynamic):
    fn = padded_session_collate({"item": PadInformation(0, max_len)}, "item", PadDirection.RIGHT, dynamic)
    return tolist(fn(samples))


positional = SequencePositionDataset(plain, pos_index, processors=[tok, TargetExtractorProcessor(features)])
pick = [0, 3, 17, len(positional) - 1, 40, 41, 7]
expected["position_batches"] = [
    {"entries": pick, "max_seq_length": m, "dynamic": d, "batch": collate([positional[i] for i in pick], m, d)}
    for m, d in ((6, False), (50, False), (50, True), (4, True))]
sess = [1, 0, 22, 5]
with_target = ItemSequenceDataset(plain, processors=[tok, TargetExtractorProcessor(features)])
expected["session_batches"] = [
    {"sessions": sess, "max_seq_length": 7, "dynamic": False, "extract_target": True,
     "batch": collate([with_target[i] for i in sess if len(dataset[i]["item"]) >= 1], 7, False)},
    {"sessions": sess, "max_seq_length": 7, "dynamic": False, "extract_target": False,
     "batch": collate([dataset[i] for i in sess], 7, False)}]
(OUT / "expected.json").write_text(json.dumps(expected, indent=1, ensure_ascii=False), encoding="utf-8")
print("sessions", len(idx), "position entries", len(pos_index), "->", OUT)
