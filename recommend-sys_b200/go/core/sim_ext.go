// sim_ext.go — PearsonBaseline is named by the north star but absent from the reference
// (core/sim.go has Cosine, MSD, Pearson only, and Sim's signature cannot see biases).  It is a
// marker value: selecting it makes KNN.Fit pass both SGD bias vectors to the device, which
// computes the cosine of the residuals r - (mu + b_left + b_right) over the co-rated entries.
package core

import "math"

// PearsonBaseline cannot be evaluated on two bare rating lists; use it through KNN only.
func PearsonBaseline(a SortedIdRatings, b SortedIdRatings) float64 { return math.NaN() }
