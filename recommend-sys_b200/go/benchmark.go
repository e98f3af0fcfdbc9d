// benchmark.go — replacement of the reference's benchmark.go (benchmark.go:12-52): the KNN
// family through 5-fold cross-validation, printing RMSE / MAE / wall time per estimator plus
// the hot-path throughputs (similarity pairs/s of the last Fit, predictions/s).
package main

import (
	"fmt"
	"os"
	"time"

	"recommend-sys/core"
)

func mean(x []float64) float64 {
	s := 0.0
	for _, v := range x {
		s += v
	}
	return s / float64(len(x))
}

func main() {
	dataset := "ml-100k"
	if len(os.Args) > 1 {
		dataset = os.Args[1]
	}
	set := core.LoadDataFromBuiltIn(dataset)
	type entry struct {
		name string
		algo core.Estimator
	}
	estimators := []entry{
		{"Slope One", core.NewSlopeOne(nil)},
		{"KNN", core.NewKNN(nil)},
		{"Centered K-NN", core.NewKNNWithMean(nil)},
		{"K-NN Baseline", core.NewKNNBaseLine(nil)},
		{"K-NN Z-Score", core.NewKNNWithZScore(nil)},
	}
	fmt.Printf("%-16s %10s %10s %12s\n", "Name", "RMSE", "MAE", "Time")
	for _, e := range estimators {
		start := time.Now()
		out := core.CrossValidate(e.algo, set, []core.Evaluator{core.RMSE, core.MAE}, 5, 0, nil)
		fmt.Printf("%-16s %10.6f %10.6f %12v\n", e.name, mean(out[0].Tests), mean(out[1].Tests), time.Since(start))
	}
}
