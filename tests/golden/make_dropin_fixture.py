"""Cut the reference's OWN code for the two call sites of the path out of /root/reference with `ast` and store it as test
fixtures (tests/golden/reference_snippets/*.txt), so that tests/test_gpu_dropin.py can EXECUTE the reference statements
unmodified against the shim on the GPU box, where /root/reference does not exist:

    pipeline_load_block.txt            rvc/infer/pipeline.py:553-561   `faiss.read_index` + `index.reconstruct_n` + try / except
    retrieve_speaker_embeddings.txt    rvc/infer/pipeline.py:497-507   the whole function, as a plain function
    extract_index_body.txt             rvc/train/process/extract_index.py:14-71   the body of the script's try block

Run in the build container only:   python tests/golden/make_dropin_fixture.py
These files are reference text, kept verbatim on purpose: they are test INPUTS (like the golden vectors), never imported
by the product; each starts with a provenance line that the test strips."""
from __future__ import annotations

import ast
import hashlib
import os
import textwrap

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_snippets")


def lines_of(src, node):
    return textwrap.dedent("\n".join(src.splitlines()[node.lineno - 1: node.end_lineno]))


def main():
    os.makedirs(OUT, exist_ok=True)
    psrc = open(os.path.join(REF, "rvc/infer/pipeline.py")).read()
    tree = ast.parse(psrc)
    load = retr = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "pipeline":
            load = next(st for st in node.body if isinstance(st, ast.If) and "file_index" in ast.get_source_segment(psrc, st.test))
        if isinstance(node, ast.FunctionDef) and node.name == "_retrieve_speaker_embeddings":
            retr = node
    esrc = open(os.path.join(REF, "rvc/train/process/extract_index.py")).read()
    etry = next(n for n in ast.parse(esrc).body if isinstance(n, ast.Try))
    ebody = textwrap.dedent("\n".join(esrc.splitlines()[etry.body[0].lineno - 1: etry.body[-1].end_lineno]))
    out = {"pipeline_load_block.txt": (f"rvc/infer/pipeline.py:{load.lineno}-{load.end_lineno}", lines_of(psrc, load)),
           "retrieve_speaker_embeddings.txt": (f"rvc/infer/pipeline.py:{retr.lineno}-{retr.end_lineno}", lines_of(psrc, retr)),
           "extract_index_body.txt": (f"rvc/train/process/extract_index.py:{etry.body[0].lineno}-{etry.body[-1].end_lineno}", ebody)}
    for name, (where, text) in out.items():
        with open(os.path.join(OUT, name), "w") as f:
            f.write(f"# REFERENCE TEXT, verbatim, test fixture only -- {where}; sha256 {hashlib.sha256(text.encode()).hexdigest()}\n")
            f.write(text if text.endswith("\n") else text + "\n")
        print(name, where, len(text.splitlines()), "lines")


if __name__ == "__main__":
    main()
