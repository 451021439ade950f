import sys, time, os
sys.path.insert(0, os.getcwd())
from tests import helpers
import tempfile, pathlib
d = pathlib.Path(tempfile.mkdtemp())
ef, cf, nf, genes = helpers.write_ex_tsv(d)
from g2vec_b200 import cli
for extra in (["-e", "5"], []):
    t = time.time()
    cli.main([ef, cf, nf, str(d / "out")] + extra)
    print("### wall %.2f s for args %s" % (time.time() - t, extra))
print(open(str(d / "out_vectors.txt")).readline()[:60])
