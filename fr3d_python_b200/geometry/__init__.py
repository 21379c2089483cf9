"""Drop-in fr3d.geometry: superpositions, discrepancy, angleofrotation, RMSD — CUDA behind."""
