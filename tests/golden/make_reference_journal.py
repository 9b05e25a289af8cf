#!/usr/bin/env python3
"""Regenerates tests/golden/reference_journal.json from the reference's shipped receipt fixture (run in the build
container, where /root/reference exists).  The receipt is a dev-mode fake, so only the journal bytes are kept."""
import json
d = json.load(open('/root/reference/data/test/test.xml-Receipt-test.json'))
out = {"source": "/root/reference/data/test/test.xml-Receipt-test.json (dev-mode fake receipt: only the journal is meaningful)",
       "inner": d["inner"], "journal_bytes": list(bytes(d['journal']['bytes']))}
json.dump(out, open(__file__.rsplit('/', 1)[0] + '/reference_journal.json', 'w'))
