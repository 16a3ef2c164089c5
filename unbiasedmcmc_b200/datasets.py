"""Real-data loaders for the reference's examples.  The data files themselves are not shipped:
pass the path of the reference's copy."""
import csv

import numpy as np

# inst/reproducegermancredit/germancredit.preparedata.R:4-9 (R's make.names of the CSV header)
_GC_CATEGORICAL = ["Account Balance", "Payment Status of Previous Credit", "Purpose", "Value Savings/Stocks",
                   "Length of current employment", "Sex & Marital Status", "Guarantors", "Most valuable available asset",
                   "Concurrent Credits", "Type of apartment", "Occupation", "Telephone", "Foreign Worker"]
_GC_QUANT = ["Duration of Credit (month)", "Credit Amount", "Instalment per cent", "Duration in Current address",
             "Age (years)", "No of Credits at this Bank", "No of dependents"]


def german_credit(csv_path):
    """german_credit.csv -> (Y, X, column names) as germancredit.preparedata.R:2-15 builds them: Y = Creditability,
    X = model.matrix(~ quantitative + factors): intercept, the 7 quantitative columns, then treatment-coded dummies
    (first level dropped, levels in increasing numeric order) of the 13 factors: n = 1000, p = 49."""
    with open(csv_path, newline="") as f:
        rows = list(csv.reader(f))
    header = [h.strip() for h in rows[0]]
    data = np.array([[float(v) for v in r] for r in rows[1:] if r], dtype=np.float64)
    col = {h: i for i, h in enumerate(header)}
    Y = data[:, col["Creditability"]].copy()
    names, cols = ["(Intercept)"], [np.ones(len(data))]
    for q in _GC_QUANT:
        names.append(q); cols.append(data[:, col[q]])
    for c in _GC_CATEGORICAL:
        v = data[:, col[c]]
        for lev in np.unique(v)[1:]:                       # contr.treatment: the first level is the baseline
            names.append(f"{c}{int(lev)}"); cols.append((v == lev).astype(np.float64))
    return Y, np.column_stack(cols), names
