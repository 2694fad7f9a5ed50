"""Generates tests/golden/disease_loci.json from the reference's simulation truth beds (sim/disease_loci_sims_minpath.bed,
sim/htt_locus.bed): the locus list (chromosome, start, stop, repeat unit) that SURVEY.md 8d config 4 is built around.
Run in the build container where /root/reference exists; the JSON travels with the repo."""
import json
import os

REF = "/root/reference/sim"
loci = []
for line in open(os.path.join(REF, "disease_loci_sims_minpath.bed")):
    f = line.split()
    if len(f) < 4:
        continue
    loci.append({"chrom": f[0], "start": int(f[1]), "stop": int(f[2]), "unit": f[3].split("_")[0], "source": "disease_loci_sims_minpath.bed"})
for line in open(os.path.join(REF, "htt_locus.bed")):
    f = line.split()
    loci.append({"chrom": f[0], "start": int(f[1]), "stop": int(f[2]), "unit": f[3], "name": f[4], "source": "htt_locus.bed"})
json.dump(loci, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "disease_loci.json"), "w"), indent=1)
print(len(loci), "loci")
