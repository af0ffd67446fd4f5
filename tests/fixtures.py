"""The reference's own fixtures (test/2seq.fastq, a1/b1, a2/b2: real MiSeq pairs), committed gzipped
under tests/golden/fixtures so that they reach the GPU box, and the golden vectors the survey recorded from the
unmodified reference on them ('Golden vectors')."""
import gzip
import hashlib
import tempfile
from pathlib import Path

GZ_DIR = Path(__file__).resolve().parent / "golden" / "fixtures"
_dir = None


def fixture_dir() -> Path:
    """directory holding the unpacked fixtures (unpacked once per process)"""
    global _dir
    if _dir is None:
        _dir = Path(tempfile.mkdtemp(prefix="nk_fixtures_"))
        for gz in sorted(GZ_DIR.glob("*.fastq.gz")):
            (_dir / gz.name[:-3]).write_bytes(gzip.decompress(gz.read_bytes()))
    return _dir


def cat_md5(cwd: Path, stem: str, k: int, norm: int, parts: int, ext="fastq"):
    h = hashlib.md5()
    for t in range(parts):
        h.update((cwd / f"{stem}.k{k}_norm{norm}_thread{t}.{ext}").read_bytes())
    return h.hexdigest()


GOLDEN = [
    # argv, (processed, printed, skipped, max used), k, norm, parts, md5 fwd, md5 rev, md5 dump
    (["-f", "a1.fastq", "-r", "b1.fastq", "-k", "15", "-d", "8", "-p", "1"], (5000, 4297, 703, 494772), 15, 8, 1,
     "3bf335a853dacaee40d90a79267e7e1a", "d747e2ba166555e4f324f04455ae5ad5", None),
    (["-f", "a1.fastq", "-r", "b1.fastq", "-k", "15", "-d", "8", "-p", "2"], (5000, 4212, 788, 494772), 15, 4, 2,
     "ba4e1e06aedc2d072aadc5907cef5a5c", "e395140bf92fc4fa8c0c930ad0b7450c", None),
    (["-f", "a1.fastq", "-r", "b1.fastq", "-k", "15", "-d", "8", "-p", "4"], (5000, 3931, 1069, 494772), 15, 2, 4,
     "553a1f70275dfb2224031192ae3cbbbc", "bdb2b6d384284d6b9ae556c1ccdc368e", None),
    (["-f", "a2.fastq", "-r", "b2.fastq", "-k", "15", "-m", "1", "-p", "8", "-d", "16", "-P"], (10000, 8013, 1987, 879127), 15, 2, 8,
     "bf47086702b9053d33db33d8c2d17b60", "79316c15e66f6c894ef28347a0d1b55d", "e040344a4e100aad2de95a36bb600ff9"),
    (["-f", "a2.fastq", "-r", "b2.fastq", "-k", "15", "-m", "1", "-p", "64", "-d", "128", "-P"], (10000, 9577, 423, 587516), 15, 2, 64,
     "dd9e4e47a02518106081e6e69076f0a3", "8bcd568c0d6a4cfc61d8e4d7611396f1", "97f1cb4e72ea80f8ac5c59381f5aef9e"),
    (["-f", "a2.fastq", "-r", "b2.fastq", "-k", "15", "-m", "1", "-p", "200", "-d", "400", "-P"], (10000, 9870, 130, 402794), 15, 2, 200,
     "2d3c561aa6c9a7cec2848423f41bce4b", "ea76b874d6f1d3e166df72e7df3bc1d6", "e07adb5ca7541798f901ac7024105711"),
    (["-f", "a2.fastq", "-r", "b2.fastq", "-k", "21", "-c", "-m", "1", "-p", "8", "-d", "16", "-P"], (10000, 6673, 3327, 685418), 21, 2, 8,
     "b4ad25dfb3f1249abe8f3d8bdfbab7df", "eef52816ce26f8bfea442ace601b8a51", "ecebb87b6ee2ded0b8b74236076fa460"),
    (["-f", "a1.fastq", "-r", "a1.fastq", "-k", "15", "-d", "8", "-p", "2"], (7445, 5608, 1837, 248238), 15, 4, 2,
     "9454638bd82210f33b380ca86235cb07", "9454638bd82210f33b380ca86235cb07", None),
    # mixed paired + single-end input lists (the second forward file has no mate)
    (["-f", "a1.fastq", "a2.fastq", "-r", "b1.fastq", "-s", "-k", "15", "-d", "4", "-p", "2"], (19948, 13957, 5991, 924786), 15, 2, 2,
     "de5059c8d6101832b462a81580232861", "090d6555b16001e594af630d7ef45212", None),
]
