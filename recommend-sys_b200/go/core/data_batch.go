// data_batch.go — the one change to core/data.go: DataSet.Predict (core/data.go:98-105)
// type-switches to an optional BatchPredictor so the whole test set is one device call.
package core

// BatchPredictor is implemented by estimators that can predict a whole DataSet at once.
type BatchPredictor interface {
	PredictBatch(userIDs, itemIDs []int) []float64
}

// Predict ratings for a set of <userId, itemId>s.
func (d *DataSet) Predict(estimator Estimator) []float64 {
	if bp, ok := estimator.(BatchPredictor); ok {
		return bp.PredictBatch(d.Users, d.Items)
	}
	predictions := make([]float64, d.Length())
	for j := 0; j < d.Length(); j++ {
		userId, itemId, _ := d.Index(j)
		predictions[j] = estimator.Predict(userId, itemId)
	}
	return predictions
}
