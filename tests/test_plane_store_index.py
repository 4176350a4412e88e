"""CPU: the plane store's index file (btb_phylo_b200/plane_store.py) -- a hint that must survive being absent,
damaged or written for another format."""
import json
import os

from btb_phylo_b200 import plane_store as ps


def test_index_round_trip_and_damage(tmp_path):
    store = str(tmp_path)
    assert ps._load_index(store) == {}
    blobs = {"ab.planes": [1000, 32, 1234, 99, "s1"], "cd.planes": [1000, 32, 1300, 100, "s2"]}
    ps._save_index(store, blobs)
    assert ps._load_index(store) == blobs
    assert [f for f in os.listdir(store) if f.endswith(".tmp")] == []          # written under a temporary name, then renamed
    open(os.path.join(store, ps.INDEX), "w").write("{not json")
    assert ps._load_index(store) == {}
    json.dump({"magic": "BTBPLN01", "blobs": blobs}, open(os.path.join(store, ps.INDEX), "w"))
    assert ps._load_index(store) == {}                                         # another format version
    json.dump({"magic": ps.MAGIC.decode(), "blobs": [1, 2]}, open(os.path.join(store, ps.INDEX), "w"))
    assert ps._load_index(store) == {}


def test_blob_header_validation(tmp_path):
    src = tmp_path / "a.fas"
    src.write_bytes(b">a\nACGT\n")
    st = os.stat(src)
    words, name = 4, b"a"
    blob = tmp_path / "x.planes"
    good = ps.HEADER.pack(ps.MAGIC, 4, words, st.st_size, st.st_mtime_ns, len(name)) + name + bytes(3 * words * 4)
    blob.write_bytes(good)
    assert ps._read_header(str(blob), st) == (4, words, "a")
    blob.write_bytes(good[:-1])
    assert ps._read_header(str(blob), st) is None                               # truncated
    blob.write_bytes(ps.HEADER.pack(ps.MAGIC, 4, words, st.st_size + 1, st.st_mtime_ns, len(name)) + name + bytes(3 * words * 4))
    assert ps._read_header(str(blob), st) is None                               # the source has another size now
    assert ps._read_header(str(tmp_path / "missing.planes"), st) is None
